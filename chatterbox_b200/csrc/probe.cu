// Hardware probe (diagnostic entry point, not on the product path): can a tcgen05 K-major SWIZZLE_128B A operand start at
// an arbitrary ROW of a tile that sits in shared memory, i.e. can one staged [rows + halo][64] tile feed every tap of a
// 1-D convolution through shifted descriptors (the "shared-memory ring for the dilated receptive field" of the HiFT
// stage kernels)?  D[128 x 64] = A[shift .. shift+127][0..63] . W[64 x 64]^T with A staged once as 160 rows.
#include "ops.h"

namespace cbx {

__global__ void __launch_bounds__(128) umma_rowshift_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmW, int shift, int mode,
                                                                  float* C) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 160 rows x 128 B
  uint8_t* sW = smem + 24576;         // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24576 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<64>(slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 160 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sW, &tmW, bar, 0, 0);
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const uint32_t a0 = smem_u32(sA) + (uint32_t)shift * 128u;
    for (int k4 = 0; k4 < 4; ++k4) {
      uint64_t da = umma_desc_sw128(a0 + k4 * 32);
      if (mode == 1) da |= (uint64_t)((a0 >> 7) & 7u) << 49;          // matrix base offset = row inside the 1024-byte swizzle atom
      umma_bf16(tmem, da, umma_desc_sw128(smem_u32(sW) + k4 * 32), umma_idesc_bf16(128, 64), k4 != 0 ? 1u : 0u);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tcgen05_fence_after();
  uint32_t r[32];
  for (int cc = 0; cc < 64; cc += 32) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + cc, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) C[(warp * 32 + lane) * 64 + cc + j] = __uint_as_float(r[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

void umma_rowshift_probe(Ctx& ctx, const __nv_bfloat16* A, const __nv_bfloat16* W, int shift, int mode, float* C) {
  CUtensorMap tmA, tmW;
  make_plane_tmap(&tmA, A, 160, 64, 160);
  make_plane_tmap(&tmW, W, 64, 64, 64);
  CBX_CHECK(cudaFuncSetAttribute(umma_rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  umma_rowshift_probe_kernel<<<1, 128, 40960, ctx.stream>>>(tmA, tmW, shift, mode, C);
  CBX_CHECK(cudaGetLastError());
}

}  // namespace cbx
