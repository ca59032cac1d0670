---
-- CUDA (B200) backend probe and FFI declarations for libluaradio_b200.so.
--
-- Drop-in next to radio/core/platform.lua: it follows the same pattern as the volk/liquid/fftw3f
-- probes there (platform.lua:277-299, kill switch :328-330).  After `require('radio_b200.platform')`:
--   platform.features.cuda  -- true when the library loaded AND a sm_100 device initialised
--   platform.libs.cuda      -- the ffi.load handle
-- Set LUARADIO_DISABLE_CUDA=1 to force the stock backends.
--
-- NOTE: written for LuaJIT 2.0/2.1; not executed in the build container (no LuaJIT there). The C ABI
-- it binds is exercised from Python ctypes (luaradio_b200/_lib.py) with the same prototypes.

local ffi = require('ffi')
local platform = require('radio.core.platform')

ffi.cdef[[
typedef struct lrb200_block_s lrb200_block_t;
typedef struct lrb200_graph_s lrb200_graph_t;

int lrb200_init(int device);
int lrb200_device_count(void);
const char *lrb200_last_error(void);
const char *lrb200_version(void);
int lrb200_sync(void);

int lrb200_block_execute(lrb200_block_t *q, const void *x, size_t n, void *y, size_t *n_out);
size_t lrb200_block_max_output(const lrb200_block_t *q, size_t n);
int lrb200_block_reset(lrb200_block_t *q);
void lrb200_block_destroy(lrb200_block_t *q);

lrb200_block_t *lrb200_fir_create_crcf(const float32_t *taps, unsigned int ntaps, unsigned int decim, unsigned int flags);
lrb200_block_t *lrb200_fir_create_cccf(const complex_float32_t *taps, unsigned int ntaps, unsigned int decim, unsigned int flags);
lrb200_block_t *lrb200_fir_create_rrrf(const float32_t *taps, unsigned int ntaps, unsigned int decim, unsigned int flags);
int lrb200_fir_set_algorithm(lrb200_block_t *q, int algo);
lrb200_block_t *lrb200_hilbert_create(const float32_t *taps, unsigned int ntaps, unsigned int flags);
lrb200_block_t *lrb200_rotator_create(double turns_per_sample, unsigned int flags);
lrb200_block_t *lrb200_discrim_create(float gain, unsigned int flags);
lrb200_block_t *lrb200_downsample_create(unsigned int factor, unsigned int elem_size, unsigned int flags);
lrb200_block_t *lrb200_iir_create_rrrf(const float32_t *b, unsigned int nb, const float32_t *a, unsigned int na, unsigned int flags);
lrb200_block_t *lrb200_iir_create_crcf(const float32_t *b, unsigned int nb, const float32_t *a, unsigned int na, unsigned int flags);
lrb200_block_t *lrb200_cmag_create(unsigned int flags);
lrb200_block_t *lrb200_c2r_create(unsigned int flags);
lrb200_block_t *lrb200_iqconv_create(const char *format, unsigned int flags);
lrb200_block_t *lrb200_realconv_create(const char *format, unsigned int flags);
lrb200_block_t *lrb200_iqsink_create(const char *format, unsigned int flags);
lrb200_block_t *lrb200_realsink_create(const char *format, unsigned int flags);

lrb200_graph_t *lrb200_graph_create(void);
int lrb200_graph_append(lrb200_graph_t *g, lrb200_block_t *q);
int lrb200_graph_commit(lrb200_graph_t *g, int fuse);
int lrb200_graph_execute(lrb200_graph_t *g, const void *x, size_t n, void *y, size_t *n_out);
size_t lrb200_graph_max_output(const lrb200_graph_t *g, size_t n);
const char *lrb200_graph_describe(const lrb200_graph_t *g);
void lrb200_graph_destroy(lrb200_graph_t *g);
]]

platform.features.cuda = false
if not os.getenv('LUARADIO_DISABLE_CUDA') then
    local available, lib = platform.load({"luaradio_b200", "libluaradio_b200.so", "libluaradio_b200.so.0"})
    if available and lib.lrb200_device_count() > 0 and lib.lrb200_init(0) == 0 then
        platform.features.cuda = true
        platform.libs.cuda = lib
    end
end

local M = {HOST = 0, DEVICE = 1}

--- Wrap a freshly created handle: NULL -> error() like firfilter.lua:199-201, GC -> destroy.
function M.own(handle, what)
    if handle == nil then
        error("Creating lrb200 " .. what .. " object: " .. ffi.string(platform.libs.cuda.lrb200_last_error()))
    end
    return ffi.gc(handle, platform.libs.cuda.lrb200_block_destroy)
end

--- Shared process() body: hand the read-only input Vector to the library, return the block-owned output.
-- `self.out` is the reusable output Vector (Vector:resize is grow-only, radio/core/vector.lua:108-136).
local n_out = ffi.new("size_t[1]")
function M.process(self, x)
    local lib = platform.libs.cuda
    local out = self.out:resize(tonumber(lib.lrb200_block_max_output(self.handle, x.length)))
    if lib.lrb200_block_execute(self.handle, x.data, x.length, out.data, n_out) ~= 0 then
        error(self.name .. ": " .. ffi.string(lib.lrb200_last_error()))
    end
    return out:resize(tonumber(n_out[0]))
end

return M
