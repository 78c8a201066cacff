// Development probe: which tensor-map / TMA variants work on this box.  Usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_global, int bytes, int x, int y, uint8_t *out)
{
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        const CUtensorMap *m = use_global ? gmap : &tmap;
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(s32(tile)), "l"((uint64_t)m), "r"(x), "r"(y), "r"(0), "r"(s32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(s32(tile)), "l"((uint64_t)m), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; !ok && spin < (1 << 22); spin++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
    if (!ok) { if (threadIdx.x == 0) out[0] = 0xEE; return; }
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    int v = argc > 1 ? atoi(argv[1]) : 0;
    int rank = (v & 1) ? 3 : 2, bw = (v & 2) ? 64 : 80, bh = 38, use_global = (v & 4) ? 1 : 0;
    int x = argc > 2 ? atoi(argv[2]) : 16, y = argc > 3 ? atoi(argv[3]) : 8;
    int W = 300, H = 200, pitch = 384;
    uint8_t *d; cudaMalloc(&d, pitch * H); std::vector<uint8_t> h(pitch * H);
    for (int i = 0; i < pitch * H; i++) h[i] = (uint8_t)((i % pitch) + 3 * (i / pitch));
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("variant %d rank %d box %dx%d global %d xy %d,%d entry %p q %d\n", v, rank, bw, bh, use_global, x, y, fn, (int)q);
    CUtensorMap m; cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, 1}, gs[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CUtensorMap *gm; cudaMalloc(&gm, sizeof(m)); cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice);
    uint8_t *out; cudaMalloc(&out, bw * bh); cudaMemset(out, 0, bw * bh);
    if (rank == 3) k<3><<<1, 128, bw * bh>>>(m, gm, use_global, bw * bh, x, y, out);
    else k<2><<<1, 128, bw * bh>>>(m, gm, use_global, bw * bh, x, y, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> o(bw * bh); cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < bh; r2++) for (int c = 0; c < bw; c++) {
            int gx = x + c, gy = y + r2; uint8_t exp = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0 : h[gy * pitch + gx];
            bad += o[r2 * bw + c] != exp;
        }
        printf("first byte %02x mismatches %d\n", o[0], bad);
    }
    return 0;
}
