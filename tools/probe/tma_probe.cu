// tma_probe.cu -- which cp.async.bulk.tensor configurations the B200 accepts for uint8 image tiles (development probe,
// one configuration per process because a rejected descriptor kills the CUDA context).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe <rank> <boxw> <boxh> <x> <y> <l2promo>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int z, int bytes, uint8_t *out) {
    __shared__ __align__(128) uint8_t s[256 * 64];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(s);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(d), "l"(&tm), "r"(x), "r"(y), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(d), "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 16) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = done ? s[i] : 0xEE;
}

int main(int argc, char **argv) {
    const int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]), l2 = atoi(argv[6]);
    const int W = 640, H = 480, B = 2, pitch = 640;
    std::vector<uint8_t> img((size_t)pitch * H * B);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d_img, *d_out;
    cudaMalloc(&d_img, img.size());
    cudaMalloc(&d_out, 65536);
    cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u}, es[3] = {1u, 1u, 1u};
    const CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d_img, dims, strides, box, es,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box %dx%d at (%d,%d) l2 %d: encode %d ", rank, bw, bh, x, y, l2, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 2; }
    const int bytes = bw * bh;
    if (rank == 2) probe<2><<<1, 128>>>(tm, x, y, 1, bytes, d_out); else probe<3><<<1, 128>>>(tm, x, y, 1, bytes, d_out);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("run: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> out(bytes);
    cudaMemcpy(out.data(), d_out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    const int frame = rank == 2 ? 0 : 1;
    for (int r2 = 0; r2 < bh; ++r2)
        for (int c = 0; c < bw; ++c) {
            const int gx = x + c, gy = y + r2;
            const uint8_t want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : img[(size_t)frame * pitch * H + (size_t)gy * pitch + gx];
            bad += out[r2 * bw + c] != want;
        }
    printf("run ok, mismatches %d of %d\n", bad, bytes);
    return bad ? 3 : 0;
}
