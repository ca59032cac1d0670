---
-- CUDA (B200) backend probe and FFI declarations for libluaradio_b200.so.
--
-- Drop-in next to radio/core/platform.lua: it follows the same pattern as the volk/liquid/fftw3f
-- probes there (platform.lua:277-299, kill switch :328-330).  After `require('radio_b200.platform')`:
--   platform.features.cuda  -- true when the library loaded AND a sm_100 device initialised
--   platform.libs.cuda      -- the ffi.load handle
-- Set LUARADIO_DISABLE_CUDA=1 to force the stock backends.
--
-- The declarations are generated from include/lrb200.h (tools/gen_lua_cdef.py -> radio_b200/cdef.lua), and
-- tests/test_lua_glue.py checks that every lib.lrb200_* call in these files is declared there.
--
-- NOTE: written for LuaJIT 2.0/2.1.  There is no LuaJIT in the build container: the logic of these files runs in the
-- CPU test-suite under a test interpreter with mock radio / ffi / library objects (tests/test_lua_exec.py); the C ABI
-- they bind is exercised from Python ctypes (luaradio_b200/_lib.py) with the same prototypes.

local ffi = require('ffi')
local platform = require('radio.core.platform')
require('radio.types')                           -- complex_float32_t / float32_t
local decl = require('radio_b200.cdef')

ffi.cdef(decl.cdef)

platform.features.cuda = false
if not os.getenv('LUARADIO_DISABLE_CUDA') then
    local available, lib = platform.load({"luaradio_b200", "libluaradio_b200.so", "libluaradio_b200.so.0"})
    if available and lib.lrb200_device_count() > 0 and lib.lrb200_init(tonumber(os.getenv('LUARADIO_CUDA_DEVICE') or 0)) == 0 then
        platform.features.cuda = true
        platform.libs.cuda = lib
    end
end

local M = {HOST = decl.constants.HOST, DEVICE = decl.constants.DEVICE,
           FIR_DIRECT = decl.constants.FIR_DIRECT, FIR_FFT = decl.constants.FIR_FFT}

--- Raise the library's error string (firfilter.lua:199-201 style).
function M.fail(what)
    error(what .. ": " .. ffi.string(platform.libs.cuda.lrb200_last_error()))
end

--- Wrap a freshly created handle: NULL -> error() like firfilter.lua:199-201, GC -> destroy.
function M.own(handle, what)
    if handle == nil then
        M.fail("Creating lrb200 " .. what .. " object")
    end
    return ffi.gc(handle, platform.libs.cuda.lrb200_block_destroy)
end

--- Give a block class its GPU form.  `create(self, flags)` makes the library handle (HOST pointers for the block's own
-- process(), DEVICE pointers when the scheduler moves the block into a flow graph); `out_type(self)` is the data
-- type of the block-owned output Vector.
function M.install(class, what, create, out_type)
    class.gpu_create = create
    class.gpu_what = what
    function class:initialize()
        -- a member of a GPU sub-graph never runs its own process(): no HOST-pointer handle.  (A subclass's initialize() --
        -- LowpassFilterBlock, SinglepoleLowpassFilterBlock ... -- has designed self.taps / b_taps by the time it calls this.)
        if self.in_gpu_subgraph then return end
        self.handle = M.own(create(self, M.HOST), what)
        self.out = out_type(self).vector()
    end
    -- the same create call with DEVICE pointers; ownership goes to the graph it is appended to
    function class:make_device_handle()
        local h = create(self, M.DEVICE)
        if h == nil then M.fail("Creating lrb200 " .. what .. " object") end
        return h
    end
end

--- Shared process() body: hand the read-only input Vector to the library, return the block-owned output.
-- `self.out` is the reusable output Vector (Vector:resize is grow-only, radio/core/vector.lua:108-136).
local n_out = ffi.new("size_t[1]")
function M.process(self, x)
    local lib = platform.libs.cuda
    local out = self.out:resize(tonumber(lib.lrb200_block_max_output(self.handle, x.length)))
    if lib.lrb200_block_execute(self.handle, x.data, x.length, out.data, n_out) ~= 0 then
        M.fail(self.name)
    end
    return out:resize(tonumber(n_out[0]))
end

--- Multi-port blocks (two inputs, or PLLBlock's two outputs).  `create(self, flags)` as above; `out_types(self)` returns
-- the array of output data types.  process(x1, x2, ...) hands one pointer per port to lrb200_block_execute_multi
-- (radio/core/block.lua:516-532 calls process() with one Vector per input, all of the same length).
function M.install_multi(class, what, create, out_types)
    class.gpu_create = create
    class.gpu_what = what
    function class:initialize()
        if self.in_gpu_subgraph then return end
        self.handle = M.own(create(self, M.HOST), what)
        self.outs = {}
        for i, t in ipairs(out_types(self)) do self.outs[i] = t.vector() end
    end
    function class:make_device_handle()
        local h = create(self, M.DEVICE)
        if h == nil then M.fail("Creating lrb200 " .. what .. " object") end
        return h
    end
end

function M.process_multi(self, ...)
    local lib = platform.libs.cuda
    local xs = {...}
    local n = xs[1].length
    local cap = tonumber(lib.lrb200_block_max_output(self.handle, n))
    local ins = ffi.new("const void*[?]", #xs)
    local outs = ffi.new("void*[?]", #self.outs)
    for i, x in ipairs(xs) do ins[i - 1] = x.data end
    for i, o in ipairs(self.outs) do outs[i - 1] = o:resize(cap).data end
    if lib.lrb200_block_execute_multi(self.handle, ins, #xs, n, outs, #self.outs, n_out) ~= 0 then
        M.fail(self.name)
    end
    for _, o in ipairs(self.outs) do o:resize(tonumber(n_out[0])) end
    return unpack(self.outs)
end

return M
