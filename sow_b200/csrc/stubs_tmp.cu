// TEMPORARY bring-up stubs (removed as the real kernels land).
#include "common.cuh"
using namespace sowb;
extern "C" {
size_t sow_merge_table_stride(void) { return 0; }
int sow_merge_grouped(const sowb_merge_entry*, int, int, void*, size_t, void*) { return set_error(SOWB_ENOTSUP, "not implemented"); }
int sow_thin_qr(const float*, int64_t, int, float*, int64_t, int, int, int, void*, size_t, void*) { return set_error(SOWB_ENOTSUP, "not implemented"); }
int tt_project(const float*, int64_t, const float*, int64_t, float*, int64_t, int, int, int, int, void*) { return set_error(SOWB_ENOTSUP, "not implemented"); }
int tt_interleave2(const void*, int, int, int, int, float*, int, void*) { return set_error(SOWB_ENOTSUP, "not implemented"); }
int tt_adam_fused2(void*, const void*, const float*, const float*, const float*, const float*, int, float*, float*, int, int, int, int, float, float, float, float, float, int, int, void*) { return set_error(SOWB_ENOTSUP, "not implemented"); }
int sow_adam_multi(void* const*, const void* const*, void* const*, void* const*, const int64_t*, int, float, float, float, float, float, float, float, int, int, void*, size_t, void*) { return set_error(SOWB_ENOTSUP, "not implemented"); }
size_t sow_adam_table_bytes(int) { return 0; }
}
