-- TEST FIXTURE: a mock `radio` namespace with the classes lua/radio_b200/blocks_patch.lua patches.  Each instantiate()
-- keeps the fields the glue's create callbacks read (self.taps, self.offset, self.gain, self.factor, self.b_taps ...).
local block = require('radio.core.block')
local types = require('radio.types')

local radio = {types = types, block = block}

local function simple(name, in_type, out_type, fields)
    local class = block.factory(name)
    function class:instantiate(...)
        local args = {...}
        for i, f in ipairs(fields or {}) do self[f] = args[i] end
        self:add_type_signature({block.Input("in", in_type)}, {block.Output("out", out_type)})
    end
    radio[name] = class
    return class
end

local C, F = types.ComplexFloat32, types.Float32
-- FIRFilterBlock(taps, use_fft, data_type): the mock takes the input type as a third argument instead of differentiating
local FIR = block.factory("FIRFilterBlock")
function FIR:instantiate(taps, use_fft, data_type)
    self.taps, self.use_fft = taps, use_fft
    self:add_type_signature({block.Input("in", data_type or C)}, {block.Output("out", data_type or C)})
end
radio.FIRFilterBlock = FIR
simple("FrequencyTranslatorBlock", C, C, {"offset"})
simple("FrequencyDiscriminatorBlock", C, F, {"gain"})
local Down = block.factory("DownsamplerBlock")
function Down:instantiate(factor, data_type)
    self.factor = factor
    self:add_type_signature({block.Input("in", data_type or F)}, {block.Output("out", data_type or F)})
end
radio.DownsamplerBlock = Down
local Up = block.factory("UpsamplerBlock")
function Up:instantiate(factor, data_type)
    self.factor = factor
    self:add_type_signature({block.Input("in", data_type or F)}, {block.Output("out", data_type or F)})
end
radio.UpsamplerBlock = Up
local IIR = block.factory("IIRFilterBlock")
function IIR:instantiate(b_taps, a_taps, data_type)
    self.b_taps, self.a_taps = b_taps, a_taps
    self:add_type_signature({block.Input("in", data_type or F)}, {block.Output("out", data_type or F)})
end
radio.IIRFilterBlock = IIR
simple("HilbertTransformBlock", F, C, {"hilbert_taps"})
simple("ComplexMagnitudeBlock", C, F)
simple("ComplexToRealBlock", C, F)
local Mul = block.factory("MultiplyConstantBlock")
function Mul:instantiate(constant, data_type)
    self.constant = constant
    self:add_type_signature({block.Input("in", data_type or F)}, {block.Output("out", data_type or F)})
end
radio.MultiplyConstantBlock = Mul

-- file blocks: only instantiate(file, format, ...) matters to the glue's constructor wrapper
for _, name in ipairs({"IQFileSource", "RealFileSource", "IQFileSink", "RealFileSink"}) do
    local class = block.factory(name)
    function class:instantiate(file, format, rate)
        self.file, self.format, self.rate = file, format .. "-table", rate
    end
    radio[name] = class
end

-- host-only blocks for the scheduler tests
simple("HostBlock", C, C)
local Source = block.factory("Source")
function Source:instantiate(data_type) self:add_type_signature({}, {block.Output("out", data_type or C)}) end
radio.Source = Source
local Sink = block.factory("Sink")
function Sink:instantiate(data_type) self:add_type_signature({block.Input("in", data_type or F)}, {}) end
radio.Sink = Sink
-- two-input element-wise blocks: the mock takes the data type as an argument instead of differentiating
for _, name in ipairs({"MultiplyBlock", "MultiplyConjugateBlock", "AddBlock", "SubtractBlock"}) do
    local class = block.factory(name)
    function class:instantiate(data_type)
        local t = data_type or C
        self:add_type_signature({block.Input("in1", t), block.Input("in2", t)}, {block.Output("out", t)})
    end
    radio[name] = class
end
-- DelayBlock keeps a host implementation (the glue falls back to it for the Bit / Byte signatures)
local Delay = block.factory("DelayBlock")
function Delay:instantiate(num_samples, data_type)
    self.num_samples = num_samples
    self:add_type_signature({block.Input("in", data_type or C)}, {block.Output("out", data_type or C)})
end
function Delay:initialize() self.host_initialized = true end
function Delay:process(x) self.host_processed = (self.host_processed or 0) + 1 return x end
radio.DelayBlock = Delay
local PLL = block.factory("PLLBlock")
function PLL:instantiate(loop_bandwidth, frequency_min, frequency_max, multiplier)
    self.loop_bw, self.freq_min, self.freq_max, self.multiplier = loop_bandwidth, frequency_min, frequency_max, multiplier or 1.0
    self:add_type_signature({block.Input("in", C)}, {block.Output("out", C), block.Output("error", F)})
end
radio.PLLBlock = PLL

-- CompositeBlock: _crawl_connections returns the stored flat map on the top-level call; start records its argument
local Composite = block.factory("CompositeBlock")
function Composite:instantiate() self.flat = {} end
function Composite:_crawl_connections(crawled_connections, composite_stack)
    self.crawl_calls = (self.crawl_calls or 0) + 1
    return crawled_connections or self.flat
end
function Composite:start(multiprocess)
    self.started_multiprocess = multiprocess
    return self
end
radio.CompositeBlock = Composite

return radio
