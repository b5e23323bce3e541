// TMA probe variants (one per process: a fault kills the context).  usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
struct Maps { alignas(64) unsigned char m[12][128]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#define PROLOGUE(BYTES)                                                                                        \
    __shared__ __align__(128) unsigned char tile[256 * 68];                                                    \
    __shared__ __align__(8) unsigned long long bar;                                                            \
    if (threadIdx.x == 0) {                                                                                    \
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");      \
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                           \
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BYTES) : "memory");
#define EPILOGUE(BYTES)                                                                                        \
    }                                                                                                          \
    __syncthreads();                                                                                           \
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory"); \
    for (int i = threadIdx.x; i < (BYTES); i += blockDim.x) out[i] = tile[i];

__global__ void k2d(const __grid_constant__ CUtensorMap map, int x, int y, int bytes, unsigned char* out) {
    PROLOGUE(bytes)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    EPILOGUE(bytes)
}
__global__ void k3d(const __grid_constant__ CUtensorMap map, int x, int y, int z, int bytes, unsigned char* out) {
    PROLOGUE(bytes)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    EPILOGUE(bytes)
}
__global__ void k3d_arr(const __grid_constant__ Maps maps, int l, int x, int y, int z, int bytes, unsigned char* out) {
    PROLOGUE(bytes)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(tile)), "l"(&maps.m[l]), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    EPILOGUE(bytes)
}
__global__ void k3d_glob(const CUtensorMap* map, int x, int y, int z, int bytes, unsigned char* out) {
    PROLOGUE(bytes)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(tile)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    EPILOGUE(bytes)
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int step = 816, rows = 518, F = 3, boxH = 44;
    const int boxW = (variant == 5) ? 128 : (variant == 6 ? 256 : 192);
    std::vector<unsigned char> h((size_t)step * rows * F);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, 256 * 68);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m;
    const int x = (variant == 7) ? 48 : 47, y = 35, z = 2;
    CUresult r;
    if (variant == 1) {
        cuuint64_t dims[2] = {step, (cuuint64_t)rows * F}; cuuint64_t str[1] = {step};
        cuuint32_t box[2] = {(cuuint32_t)boxW, boxH}, es[2] = {1, 1};
        r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[3] = {step, rows, F}; cuuint64_t str[2] = {step, (cuuint64_t)step * rows};
        cuuint32_t box[3] = {(cuuint32_t)boxW, boxH, 1}, es[3] = {1, 1, 1};
        r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r) { printf("variant %d encode failed %d\n", variant, (int)r); return 1; }
    Maps maps; memset(&maps, 0, sizeof maps);
    for (int l = 0; l < 12; l++) memcpy(maps.m[l], &m, 128);
    const int bytes = boxW * boxH;
    if (variant == 1) k2d<<<1, 128>>>(m, x, z * rows + y, bytes, o);
    else if (variant == 2 || variant >= 5) k3d<<<1, 128>>>(m, x, y, z, bytes, o);
    else if (variant == 3) k3d_arr<<<1, 128>>>(maps, 5, x, y, z, bytes, o);
    else if (variant == 4) { CUtensorMap* dm; cudaMalloc(&dm, 128); cudaMemcpy(dm, &m, 128, cudaMemcpyHostToDevice); k3d_glob<<<1, 128>>>(dm, x, y, z, bytes, o); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d kernel: %s\n", variant, cudaGetErrorString(e));
    if (e) return 1;
    std::vector<unsigned char> res(bytes);
    cudaMemcpy(res.data(), o, res.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < boxH; yy++) for (int xx = 0; xx < boxW; xx++) {
        unsigned char want = (x + xx < step) ? h[((size_t)z * rows + y + yy) * step + x + xx] : 0;
        if (res[yy * boxW + xx] != want) bad++;
    }
    printf("variant %d mismatches: %d\n", variant, bad);
    return bad != 0;
}
