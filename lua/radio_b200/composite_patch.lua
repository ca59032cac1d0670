---
-- GPU scheduler hook for radio/core/composite.lua.  A CUDA context does not survive fork(), and
-- _initialize() runs in the parent before the fork loop (composite.lua:443 then :568-636), so when the CUDA
-- backend is active the top block runs through the single-process path (composite.lua:647-707) and
-- maximal linear runs of GPU blocks are collapsed into ONE lrb200 flow graph: one process() call per source
-- vector, device-resident intermediates, fused kernels, H2D/D2H only at the ends of the run.
--
--   local top = radio.CompositeBlock(); top:connect(...); top:run()   -- unchanged user code

local ffi = require('ffi')
local platform = require('radio.core.platform')
local block = require('radio.core.block')
local b200 = require('radio_b200.platform')

--- A block that stands for a run of connected GPU blocks [first .. last].
local GPUChainBlock = block.factory("GPUChainBlock")

function GPUChainBlock:instantiate(blocks)
    self.blocks = blocks
    self:add_type_signature({block.Input("in", blocks[1]:get_input_type())},
                            {block.Output("out", blocks[#blocks]:get_output_type())})
end

function GPUChainBlock:get_rate()
    return self.blocks[#self.blocks]:get_rate()
end

function GPUChainBlock:initialize()
    local lib = platform.libs.cuda
    self.graph = ffi.gc(lib.lrb200_graph_create(), lib.lrb200_graph_destroy)
    for _, b in ipairs(self.blocks) do
        -- each GPU block exposes make_device_handle(): the same create call as initialize() with b200.DEVICE
        local h = b:make_device_handle()
        if lib.lrb200_graph_append(self.graph, h) ~= 0 then
            error("graph_append: " .. ffi.string(lib.lrb200_last_error()))
        end
        ffi.gc(h, nil)      -- ownership moved to the graph
    end
    if lib.lrb200_graph_commit(self.graph, 1) ~= 0 then
        error("graph_commit: " .. ffi.string(lib.lrb200_last_error()))
    end
    self.out = self:get_output_type().vector()
end

local n_out = ffi.new("size_t[1]")
function GPUChainBlock:process(x)
    local lib = platform.libs.cuda
    local out = self.out:resize(tonumber(lib.lrb200_graph_max_output(self.graph, x.length)))
    if lib.lrb200_graph_execute(self.graph, x.data, x.length, out.data, n_out) ~= 0 then
        error("graph_execute: " .. ffi.string(lib.lrb200_last_error()))
    end
    return out:resize(tonumber(n_out[0]))
end

return {GPUChainBlock = GPUChainBlock}
