// Probe: which TMA box shapes does sm_100a accept for a planar fp32 image [4][H][W]?  usage: tma_probe BOXW BOXH X Y [rank]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, float *out, int n, int x, int y, int z, int rank) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 4) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(d), "l"(&tmap), "r"(b), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d), "l"(&tmap), "r"(b), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ((float *)smem)[i];
}
int main(int argc, char **argv) {
    int bw = atoi(argv[1]), bh = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), rank = argc > 5 ? atoi(argv[5]) : 3;
    const int W = 512, H = 256;
    std::vector<float> h((size_t)W * H * 4);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int n = bw * bh;
    cudaMalloc(&o, n * 4);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t dims[3] = {W, H, 4}, strides[2] = {W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((PFN_encodeTiled)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %dx%d rank %d: encode failed %d\n", bw, bh, rank, (int)r); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    k<<<1, 128, (n * 4 + 127) / 128 * 128, 0>>>(map, o, n, x, y, 1, rank);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> g(n);
    cudaMemcpy(g.data(), o, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < bh && e == cudaSuccess; j++)
        for (int i = 0; i < bw; i++) {
            float want = (x + i < W && y + j < H) ? h[(size_t)(rank == 3 ? 1 : 0) * W * H + (size_t)(y + j) * W + x + i] : 0.0f;
            if (g[j * bw + i] != want) bad++;
        }
    printf("box %dx%d at (%d,%d) rank %d: %s, %d wrong values\n", bw, bh, x, y, rank, cudaGetErrorString(e), bad);
    return 0;
}
