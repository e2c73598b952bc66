// Probe: which 2-D TMA load configurations work on this part.  usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../../imagestitch_b200/csrc/tma.cuh"
using namespace is;

template <int BYTES>
__global__ void k_gc(const __grid_constant__ CUtensorMap map, int c0, int c1, unsigned char* out) {
    __shared__ __align__(128) unsigned char buf[BYTES];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, BYTES); tma_load_2d(buf, &map, c0, c1, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BYTES; i += blockDim.x) out[i] = buf[i];
}
template <int BYTES>
__global__ void k_gm(const CUtensorMap* map, int c0, int c1, unsigned char* out, int fence) {
    __shared__ __align__(128) unsigned char buf[BYTES];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { if (fence) tensormap_acquire(map); mbar_expect_tx(&bar, BYTES); tma_load_2d(buf, map, c0, c1, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BYTES; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    // source: 64 rows x 1056 int16 (pitch 2112 B), value = row * 1056 + col
    const int W = 1056, H = 64;
    std::vector<int16_t> h((size_t)W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (int16_t)(i & 0x7fff);
    int16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap m;
    const bool u8 = (v / 10) == 1;     // variants 1x: uint8 view, box 192 x 32
    cuuint64_t dims[2] = {(cuuint64_t)(u8 ? W * 2 : W), (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W * 2};
    cuuint32_t box[2] = {(cuuint32_t)(u8 ? 208 : 112), (cuuint32_t)(u8 ? 32 : 18)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        (v % 10) >= 6 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode=%d\n", v, (int)r);
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    unsigned char* out; cudaMalloc(&out, 8192); cudaMemset(out, 0xee, 8192);
    const int sub = v % 10;
    // sub 0: grid_constant, aligned coords (8, 2); 1: grid_constant, unaligned (573, 2); 2: grid_constant, negative (-3, -1)
    // sub 3: global desc + fence, aligned; 4: global desc + fence, unaligned; 5: global desc, no fence, aligned; 6: as 1 with L2 promotion none
    int c0 = 8, c1 = 2;
    if (sub == 1 || sub == 4 || sub == 6) c0 = 573;
    if (sub == 2) { c0 = -3; c1 = -1; }
    if (sub == 7) { c0 = -8; c1 = -1; }      // negative but 16-byte aligned
    if (sub == 8) { c0 = 1048; c1 = 60; }    // runs off the right / bottom edge, aligned
    if (u8) c0 *= 2;                          // same byte offsets for the uint8 view
    if (u8) { if (sub <= 2 || sub == 6 || sub == 7 || sub == 8) k_gc<6656><<<1, 128>>>(m, c0, c1, out); else k_gm<6656><<<1, 128>>>(dm, c0, c1, out, sub != 5); }
    else { if (sub <= 2 || sub == 6 || sub == 7 || sub == 8) k_gc<4032><<<1, 128>>>(m, c0, c1, out); else k_gm<4032><<<1, 128>>>(dm, c0, c1, out, sub != 5); }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d sync=%s\n", v, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<unsigned char> o(8192); cudaMemcpy(o.data(), out, 8192, cudaMemcpyDeviceToHost);
        const int16_t* s = reinterpret_cast<const int16_t*>(o.data());
        if (!u8) printf("  first elems: %d %d %d %d (expect from row %d col %d: %d)\n", s[0], s[1], s[2], s[3], c1, c0, (c1 >= 0 && c0 >= 0) ? ((c1 * W + c0) & 0x7fff) : 0);
        else printf("  first bytes: %d %d %d %d\n", o[0], o[1], o[2], o[3]);
        if (!u8 && sub == 7) printf("  row 1 (= source row 0) elems 8..11: %d %d %d %d (expect 0 1 2 3), elems 0..1: %d %d (expect 0 0)\n", s[112 + 8], s[112 + 9], s[112 + 10], s[112 + 11], s[112], s[113]);
        if (!u8 && sub == 8) printf("  row 0 elems 0..1: %d %d (expect %d %d), elems 8..9: %d %d (expect 0 0 beyond the row end), row 4 elem 0: %d (expect 0)\n", s[0], s[1], (60 * 1056 + 1048) & 0x7fff, (60 * 1056 + 1049) & 0x7fff, s[8], s[9], s[4 * 112]);
    }
    return 0;
}
