// tools/membench.cu -- HBM access-order micro-benchmark (experiment, not product): how does the order in which a
// persistent grid sweeps memory change achieved DRAM bandwidth for write-only / read-only / copy traffic?
//   mode 0 linear      : block b handles chunk b, b+grid, ... of 8 KB chunks, grid = total chunks (one chunk per block)
//   mode 1 round-robin : persistent grid of G blocks, block b handles chunks b, b+G, b+2G, ... (compact moving window)
//   mode 2 streams     : persistent grid of G blocks, block b handles its own contiguous 1/G of the buffer
//   mode 3 dynamic     : persistent grid of G blocks, each block takes the next chunk from an atomic counter
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int OP>  // 0 write, 1 read, 2 copy
__global__ void __launch_bounds__(256) k(const uint4 *__restrict__ in, uint4 *__restrict__ out, long long n_chunks,
                                         int chunk_vec, int mode, uint4 *sink)
{
    const long long G = gridDim.x;
    long long c0, c1, step;
    if (mode == 2) { long long per = (n_chunks + G - 1) / G; c0 = blockIdx.x * per; c1 = min(n_chunks, c0 + per); step = 1; }
    else { c0 = blockIdx.x; c1 = n_chunks; step = G; }
    uint4 acc = make_uint4(0, 0, 0, 0);
    __shared__ long long next;
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>(sink) + 4;
    for (long long c = c0; c < c1; c += step) {
        if (mode == 3) {
            __syncthreads();
            if (threadIdx.x == 0) next = (long long)atomicAdd(ctr, 1ull);
            __syncthreads();
            c = next;
            if (c >= n_chunks) break;
        }
        const long long base = c * chunk_vec;
        for (int i = threadIdx.x; i < chunk_vec; i += blockDim.x * 4) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int idx = i + j * blockDim.x;
                if (OP >= 1) { if (idx < chunk_vec) v[j] = __ldg(in + base + idx); }
                else v[j] = make_uint4(idx, 1, 2, 3);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int idx = i + j * blockDim.x;
                if (idx < chunk_vec) {
                    if (OP == 1) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z ^= v[j].z; acc.w ^= v[j].w; }
                    else out[base + idx] = v[j];
                }
            }
        }
    }
    if (OP == 1 && acc.x == 0x12345678u && acc.y == 1u) *sink = acc;
}

int main(int argc, char **argv)
{
    const size_t bytes = 1152000000;  // one direction of the 5000 x 320x240 RGB stream
    uint4 *a, *b, *sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *opn[] = {"write", "read", "copy"};
    const char *mn[] = {"linear", "round-robin", "streams", "dynamic"};
    for (int chunk_kb : {8, 32}) {
        const int chunk_vec = chunk_kb * 1024 / 16;
        const long long n_chunks = bytes / (chunk_kb * 1024);
        for (int op = 0; op < 3; op++)
            for (int mode = 0; mode < 4; mode++)
                for (int threads : {128, 256})
                    for (int G : {148, 296, 444, 888, 1776}) {
                        if (mode == 0 && G != 148) continue;
                        const int grid = mode == 0 ? (int)n_chunks : G;
                        float best = 1e9;
                        for (int it = 0; it < 6; it++) {
                            if (mode == 3) CK(cudaMemset(sink, 0, 64));
                            cudaEventRecord(e0);
                            if (op == 0) k<0><<<grid, threads>>>(a, b, n_chunks, chunk_vec, mode, sink);
                            if (op == 1) k<1><<<grid, threads>>>(a, b, n_chunks, chunk_vec, mode, sink);
                            if (op == 2) k<2><<<grid, threads>>>(a, b, n_chunks, chunk_vec, mode, sink);
                            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                            float ms; cudaEventElapsedTime(&ms, e0, e1);
                            if (it > 0 && ms < best) best = ms;
                        }
                        const double gb = (op == 2 ? 2.0 : 1.0) * n_chunks * chunk_kb * 1024 / 1e9;
                        printf("chunk %3d KB  %-5s %-11s threads %3d grid %7d : %.4f ms  %7.1f GB/s\n", chunk_kb, opn[op],
                               mn[mode], threads, grid, best, gb / best * 1e3);
                    }
    }
    return 0;
}
