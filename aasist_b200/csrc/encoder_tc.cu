// placeholder until the tcgen05 path lands
#include "tc.cuh"
namespace aasist {
int tc_finalize(aasist_handle*) { set_error("f16x3 tensor-core path not built"); return AASIST_E_STATE; }
void tc_destroy(aasist_handle*) {}
size_t tc_workspace_bytes(const aasist_handle*, int, int) { return 0; }
int tc_encode(aasist_handle*, const float*, int, int, float**, void*, cudaStream_t) { set_error("f16x3 path not built"); return AASIST_E_STATE; }
int tc_frontend_to_f32(aasist_handle*, const float*, int, int, float*, void*, int64_t, cudaStream_t) { set_error("f16x3 path not built"); return AASIST_E_STATE; }
int tc_block_f32io(aasist_handle*, int, int, const float*, int, int, float*, void*, int64_t, cudaStream_t) { set_error("f16x3 path not built"); return AASIST_E_STATE; }
}
