---
-- The remaining hot-path blocks: each function below is the `if platform.features.cuda then` branch for
-- the file named in its comment.  Apply with  require('radio_b200.blocks_patch')(require('radio')).

local math = require('math')
local platform = require('radio.core.platform')
local types = require('radio.types')
local b200 = require('radio_b200.platform')

return function (radio)
    if not platform.features.cuda then return end
    local lib = platform.libs.cuda

    require('radio_b200.firfilter_patch')(radio.FIRFilterBlock)

    -- radio/blocks/signal/frequencytranslator.lua:32 (before the volk branch)
    function radio.FrequencyTranslatorBlock:initialize()
        self.handle = b200.own(lib.lrb200_rotator_create(self.offset / self:get_rate(), b200.HOST), "rotator")
        self.out = types.ComplexFloat32.vector()
    end
    radio.FrequencyTranslatorBlock.process = b200.process

    -- radio/blocks/signal/frequencydiscriminator.lua:33-40
    function radio.FrequencyDiscriminatorBlock:initialize()
        self.handle = b200.own(lib.lrb200_discrim_create(self.gain, b200.HOST), "discriminator")
        self.out = types.Float32.vector()
    end
    radio.FrequencyDiscriminatorBlock.process = b200.process

    -- radio/blocks/signal/downsampler.lua:40-56
    function radio.DownsamplerBlock:initialize()
        local data_type = self:get_input_type()
        self.handle = b200.own(lib.lrb200_downsample_create(self.factor, data_type == types.ComplexFloat32 and 8 or 4, b200.HOST), "downsampler")
        self.out = data_type.vector()
    end
    radio.DownsamplerBlock.process = b200.process

    -- radio/blocks/signal/iirfilter.lua:63 (before the liquid branch); Singlepole*/FMDeemphasis inherit it
    function radio.IIRFilterBlock:initialize()
        local data_type = self:get_input_type()
        local create = data_type == types.ComplexFloat32 and lib.lrb200_iir_create_crcf or lib.lrb200_iir_create_rrrf
        self.handle = b200.own(create(self.b_taps.data, self.b_taps.length, self.a_taps.data, self.a_taps.length, b200.HOST), "iir")
        self.out = data_type.vector()
    end
    radio.IIRFilterBlock.process_complex = b200.process
    radio.IIRFilterBlock.process_real = b200.process

    -- radio/blocks/signal/hilberttransform.lua:39
    function radio.HilbertTransformBlock:initialize()
        self.handle = b200.own(lib.lrb200_hilbert_create(self.hilbert_taps.data, self.hilbert_taps.length, b200.HOST), "hilbert")
        self.out = types.ComplexFloat32.vector()
    end
    radio.HilbertTransformBlock.process = b200.process

    -- radio/blocks/signal/complexmagnitude.lua:24-36, complextoreal.lua:23-35
    function radio.ComplexMagnitudeBlock:initialize()
        self.handle = b200.own(lib.lrb200_cmag_create(b200.HOST), "cmag")
        self.out = types.Float32.vector()
    end
    radio.ComplexMagnitudeBlock.process = b200.process
    function radio.ComplexToRealBlock:initialize()
        self.handle = b200.own(lib.lrb200_c2r_create(b200.HOST), "c2r")
        self.out = types.Float32.vector()
    end
    radio.ComplexToRealBlock.process = b200.process
end
