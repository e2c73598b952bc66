// Minimal reproducer of the ptxas miscompile met in k_orb_fast (CUDA 12.9.86, -arch sm_100a, -O3): the FAST corner score written as
//     best = max over the 16 arcs of nine circle pixels of max(min d, -(max d))
// comes out as if the negation were missing from the second arc on (the result equals max over arcs of max(min d, max d)).
// The PTX is right (16 neg.s32); in the SASS one IMAD.MOV -R survives and the other maxima are folded into VIMNMX3 without it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ptxas_vimnmx3_neg scripts/probe/ptxas_vimnmx3_neg.cu && ./ptxas_vimnmx3_neg
//   cuobjdump -sass ptxas_vimnmx3_neg | grep -c "IMAD.MOV R[0-9]*, RZ, RZ, -R"      # 1 where 16 negations are needed
//
// Prints the number of inputs on which the device result differs from the host's; 0 on a correct compiler.
#include <cstdio>
#include <cstdlib>
#include <vector>

__host__ __device__ inline int score_of(const int* d) {
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) {
            const int v = d[(k + j) & 15];
            mn = mn < v ? mn : v;
            mx = mx > v ? mx : v;
        }
        const int a = mn > -mx ? mn : -mx;
        best = best > a ? best : a;
    }
    return best;
}

__global__ void k_score(const int* __restrict__ in, int* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = in[16 * i + k];
    out[i] = score_of(d);
}

int main() {
    const int n = 1 << 16;
    std::vector<int> h(16 * n), want(n), got(n);
    srand(1);
    for (int& v : h) v = rand() % 511 - 255;
    for (int i = 0; i < n; ++i) want[i] = score_of(&h[16 * i]);
    int *din = nullptr, *dout = nullptr;
    if (cudaMalloc(&din, h.size() * sizeof(int)) != cudaSuccess || cudaMalloc(&dout, n * sizeof(int)) != cudaSuccess) { std::printf("no device\n"); return 2; }
    cudaMemcpy(din, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice);
    k_score<<<(n + 255) / 256, 256>>>(din, dout, n);
    cudaMemcpy(got.data(), dout, n * sizeof(int), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += got[i] != want[i];
    std::printf("%d of %d scores differ from the host's\n", bad, n);
    return bad ? 1 : 0;
}
