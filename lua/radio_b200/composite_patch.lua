---
-- GPU scheduler hook for radio/core/composite.lua.  A CUDA context does not survive fork(), and
-- _initialize() runs in the parent before the fork loop (composite.lua:443 then :568-636), so when the CUDA
-- backend is active the top block runs through the single-process path (composite.lua:647-707) and every
-- MAXIMAL LINEAR RUN of GPU blocks in the crawled connection graph is collapsed into ONE lrb200 flow graph:
-- one process() call per source vector, device-resident intermediates, fused kernels, H2D/D2H only at the
-- two ends of the run.  CPU blocks, multi-input blocks and fan-out points stay ordinary blocks at the edges.
--
--   local top = radio.CompositeBlock(); top:connect(...); top:run()   -- unchanged user code
--
-- install(radio) wraps two methods of CompositeBlock:
--   _crawl_connections (composite.lua:343-381)  the top-level call returns the flat input-port -> output-port map;
--                                               GPU runs are substituted there, before _connect_pipes (:383-393)
--   start (composite.lua:534-545)               forces multiprocess = false
-- Set LUARADIO_B200_SUPERCHUNK=<samples> to pack the per-vector calls into pinned super-chunks
-- (lrb200_graph_set_superchunk): throughput of the reference's 8192-sample vectors goes from launch-bound to
-- memcpy-bound; outputs then arrive in bursts, and the last partial super-chunk is only pushed out by cleanup().

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
    if self.graph == nil then b200.fail("graph_create") end
    for _, b in ipairs(self.blocks) do
        -- make_device_handle(): the block's own create call with DEVICE pointers (radio_b200/platform.lua install)
        local h = b:make_device_handle()
        if lib.lrb200_graph_append(self.graph, h) ~= 0 then b200.fail("graph_append") end
    end
    if lib.lrb200_graph_commit(self.graph, 1) ~= 0 then b200.fail("graph_commit") end
    local sc = tonumber(os.getenv('LUARADIO_B200_SUPERCHUNK') or 0)
    if sc > 0 and lib.lrb200_graph_set_superchunk(self.graph, sc) ~= 0 then b200.fail("graph_set_superchunk") end
    self.out = self:get_output_type().vector()
end

local n_out = ffi.new("size_t[1]")
function GPUChainBlock:process(x)
    local lib = platform.libs.cuda
    local out = self.out:resize(tonumber(lib.lrb200_graph_max_output(self.graph, x.length)))
    if lib.lrb200_graph_execute(self.graph, x.data, x.length, out.data, n_out) ~= 0 then b200.fail("graph_execute") end
    return out:resize(tonumber(n_out[0]))
end

function GPUChainBlock:cleanup()
    -- super-chunk mode: hand the pending samples to the downstream pipes before they are closed
    local lib = platform.libs.cuda
    local out = self.out:resize(tonumber(lib.lrb200_graph_max_output(self.graph, 0)))
    if lib.lrb200_graph_flush(self.graph, out.data, n_out) ~= 0 then b200.fail("graph_flush") end
    out:resize(tonumber(n_out[0]))
    if out.length > 0 then
        for _, p in ipairs(self.outputs[1].pipes) do p:write(out) end
    end
end

local M = {GPUChainBlock = GPUChainBlock}

--- Is `b` a concrete GPU block that can sit inside a linear run?
local function is_gpu(b)
    return b.make_device_handle ~= nil and #b.inputs == 1 and #b.outputs == 1
end

--- Substitute every maximal linear run (length >= 2) of GPU blocks in the crawled connection map.
-- connections: InputPort -> OutputPort over concrete blocks (composite.lua:343-381).
function M.collapse_gpu_runs(connections)
    -- consumers per output port
    local consumers = {}
    for input, output in pairs(connections) do
        consumers[output] = consumers[output] or {}
        table.insert(consumers[output], input)
    end
    local function next_in_run(b)          -- the single GPU consumer of b's output, if the edge is 1:1
        local c = consumers[b.outputs[1]]
        if c and #c == 1 and is_gpu(c[1].owner) then return c[1].owner end
        return nil
    end
    local function prev_in_run(b)
        local up = connections[b.inputs[1]]
        if up and is_gpu(up.owner) and next_in_run(up.owner) == b then return up.owner end
        return nil
    end
    -- plan first, rewrite afterwards: adding keys to a table while pairs() walks it is undefined in Lua
    local seen, runs = {}, {}
    for input, _ in pairs(connections) do
        local b = input.owner
        if is_gpu(b) and not seen[b] and prev_in_run(b) == nil then
            local run = {b}
            seen[b] = true
            local nb = next_in_run(b)
            while nb do
                run[#run + 1] = nb
                seen[nb] = true
                nb = next_in_run(nb)
            end
            if #run >= 2 then runs[#runs + 1] = run end
        end
    end
    for _, run in ipairs(runs) do
        local chain = GPUChainBlock(run)
        chain:differentiate({run[1]:get_input_type()})
        local first_in, last_out = run[1].inputs[1], run[#run].outputs[1]
        -- upstream edge now ends at the chain's input; edges inside the run disappear
        connections[chain.inputs[1]] = connections[first_in]
        for _, rb in ipairs(run) do connections[rb.inputs[1]] = nil end
        -- downstream consumers now read the chain's output
        for _, cin in ipairs(consumers[last_out] or {}) do connections[cin] = chain.outputs[1] end
    end
    return connections
end

function M.install(radio)
    if not platform.features.cuda then return end
    local CompositeBlock = radio.CompositeBlock
    local crawl, start = CompositeBlock._crawl_connections, CompositeBlock.start
    function CompositeBlock:_crawl_connections(crawled_connections, composite_stack)
        local top_level = crawled_connections == nil
        local connections = crawl(self, crawled_connections, composite_stack)
        if top_level then M.collapse_gpu_runs(connections) end
        return connections
    end
    function CompositeBlock:start(multiprocess)
        return start(self, false)       -- one process: the CUDA context cannot be forked (composite.lua:568-636)
    end
end

return M
