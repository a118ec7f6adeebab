// placeholder -- replaced by the tcgen05 implementation
#include "common.cuh"
namespace pnr {
size_t mlp_tc_packed_bytes(const pnr_mlp& m) { return 256; }
int mlp_tc_pack(const pnr_mlp& m, void* dst, size_t dst_bytes, cudaStream_t st) { set_err("tc path not built"); return PNR_ERR_UNSUPPORTED; }
size_t net_tc_workspace(const pnr_scene& sc, const pnr_mlp& m, int SB, long long P) { return 256; }
int net_forward_tc(const pnr_scene& sc, const pnr_mlp& m, const float* xyz, const float* viewdirs, const float* rays, const float* z, int K, int SB, long long P, float* out, void* ws, size_t ws_bytes, cudaStream_t st) { set_err("tc path not built"); return PNR_ERR_UNSUPPORTED; }
int mlp_forward_tc_rows(const pnr_mlp& m, const float* zx, int SB, int NS, int P, float* out, void* ws, size_t ws_bytes, cudaStream_t st) { set_err("tc path not built"); return PNR_ERR_UNSUPPORTED; }
size_t mlp_tc_rows_workspace(const pnr_mlp& m, int SB, int NS, int P) { return 256; }
}
