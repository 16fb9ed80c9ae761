// nvlink_probe.cu -- what store pattern does NVLink like?  Single process, devices 0 and 1 with peer access.
// Measures, for a 512 MB buffer, the bandwidth of: cudaMemcpyPeerAsync, a kernel writing the peer with 8-byte /
// 16-byte per-lane stores in 256 B / 512 B warp rows (rows scattered with a large stride like the fused remap
// pass), and TMA bulk stores shared -> peer global of 256 B .. 16 KB pieces.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp writes rows of ROWB bytes; consecutive rows of a warp are `stride` bytes apart in the destination
template <int VEC>
__global__ void k_store(double *dst, const double *src, long long nrows, long long row_elems, long long stride_elems) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < nrows; r += nwarps) {
        // scatter: row r goes to slot (r % 4096) * stride + (r / 4096) * row_elems
        const long long d = (r % 4096) * stride_elems + (r / 4096) * row_elems;
        const long long s = r * row_elems;
        if (VEC == 1) __stcs(dst + d + lane, src[s + lane]);
        else {
            const double2 v = *reinterpret_cast<const double2 *>(src + s + 2 * lane);
            __stcs(reinterpret_cast<double2 *>(dst + d + 2 * lane), v);
        }
    }
}
// TMA bulk store: every CTA stages PIECE bytes in shared memory once, then streams them to the destination
__global__ void k_bulk(double *dst, long long npieces, int piece_bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < piece_bytes / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = (double)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        int inflight = 0;
        for (long long p = blockIdx.x; p < npieces; p += gridDim.x) {
            char *g = reinterpret_cast<char *>(dst) + p * (long long)piece_bytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(sm)), "r"(piece_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t bytes = 512ull << 20;
    double *src, *loc, *peer;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes)); CK(cudaMemset(peer, 0, bytes));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&loc, bytes)); CK(cudaMemset(src, 1, bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    auto report = [&](const char *what, float ms_) { printf("%-58s %8.3f ms %8.1f GB/s\n", what, ms_, bytes / ms_ / 1e6); };
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a); CK(cudaMemcpyPeerAsync(peer, 1, src, 0, bytes, 0)); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    report("cudaMemcpyPeerAsync 0 -> 1", ms);
    for (int target = 0; target < 2; ++target) {
        double *dst = target ? peer : loc;
        const char *tn = target ? "peer " : "local";
        char name[128];
        for (int vec = 1; vec <= 2; ++vec) {
            const long long row_elems = 32 * vec, nrows = bytes / 8 / row_elems, stride = nrows / 4096 * row_elems;
            for (int blocks : {148 * 4, 148 * 16}) {
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(a);
                    if (vec == 1) k_store<1><<<blocks, 256>>>(dst, src, nrows, row_elems, stride);
                    else k_store<2><<<blocks, 256>>>(dst, src, nrows, row_elems, stride);
                    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
                }
                CK(cudaGetLastError());
                snprintf(name, sizeof(name), "%s st.cs %2d B/lane, %3d B rows scattered, %5d CTAs", tn, 8 * vec, 256 * vec, blocks);
                report(name, ms);
            }
        }
        for (int piece : {256, 512, 2048, 16384}) {
            const long long np = bytes / piece;
            for (int blocks : {148, 148 * 8}) {
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(a);
                    k_bulk<<<blocks, 128, piece>>>(dst, np, piece);
                    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
                }
                CK(cudaGetLastError());
                snprintf(name, sizeof(name), "%s TMA bulk store %5d B pieces, %5d CTAs", tn, piece, blocks);
                report(name, ms);
            }
        }
    }
        // ---- both directions at once (the all-to-all is bidirectional): device 0 -> 1 and 1 -> 0 concurrently ----
    {
        double *src1, *peer0;
        CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&src1, bytes)); CK(cudaMemset(src1, 1, bytes));
        CK(cudaSetDevice(0)); CK(cudaMalloc(&peer0, bytes));
        cudaEvent_t a1, b1; CK(cudaSetDevice(1)); cudaEventCreate(&a1); cudaEventCreate(&b1); CK(cudaSetDevice(0));
        for (int vec = 1; vec <= 2; ++vec) {
            const long long row_elems = 32 * vec, nrows = bytes / 8 / row_elems, stride = nrows / 4096 * row_elems;
            float ms0 = 0, ms1 = 0;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaSetDevice(0)); cudaEventRecord(a);
                if (vec == 1) k_store<1><<<148 * 16, 256>>>(peer, src, nrows, row_elems, stride);
                else k_store<2><<<148 * 16, 256>>>(peer, src, nrows, row_elems, stride);
                cudaEventRecord(b);
                CK(cudaSetDevice(1)); cudaEventRecord(a1);
                if (vec == 1) k_store<1><<<148 * 16, 256>>>(peer0, src1, nrows, row_elems, stride);
                else k_store<2><<<148 * 16, 256>>>(peer0, src1, nrows, row_elems, stride);
                cudaEventRecord(b1);
                CK(cudaSetDevice(0)); cudaEventSynchronize(b); cudaEventElapsedTime(&ms0, a, b);
                CK(cudaSetDevice(1)); cudaEventSynchronize(b1); cudaEventElapsedTime(&ms1, a1, b1);
                CK(cudaSetDevice(0));
            }
            char name[128];
            snprintf(name, sizeof(name), "bidirectional st.cs %2d B/lane %3d B rows: dev0 / dev1", 8 * vec, 256 * vec);
            report(name, ms0); report(name, ms1);
        }
        {
            float ms0 = 0, ms1 = 0;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaSetDevice(0)); cudaEventRecord(a); CK(cudaMemcpyPeerAsync(peer, 1, src, 0, bytes, 0)); cudaEventRecord(b);
                CK(cudaSetDevice(1)); cudaEventRecord(a1); CK(cudaMemcpyPeerAsync(peer0, 0, src1, 1, bytes, 0)); cudaEventRecord(b1);
                CK(cudaSetDevice(0)); cudaEventSynchronize(b); cudaEventElapsedTime(&ms0, a, b);
                CK(cudaSetDevice(1)); cudaEventSynchronize(b1); cudaEventElapsedTime(&ms1, a1, b1);
                CK(cudaSetDevice(0));
            }
            report("bidirectional cudaMemcpyPeerAsync: dev0", ms0); report("bidirectional cudaMemcpyPeerAsync: dev1", ms1);
        }
        for (int piece : {512, 16384}) {
            const long long np = bytes / piece;
            float ms0 = 0, ms1 = 0;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaSetDevice(0)); cudaEventRecord(a); k_bulk<<<148 * 8, 128, piece>>>(peer, np, piece); cudaEventRecord(b);
                CK(cudaSetDevice(1)); cudaEventRecord(a1); k_bulk<<<148 * 8, 128, piece>>>(peer0, np, piece); cudaEventRecord(b1);
                CK(cudaSetDevice(0)); cudaEventSynchronize(b); cudaEventElapsedTime(&ms0, a, b);
                CK(cudaSetDevice(1)); cudaEventSynchronize(b1); cudaEventElapsedTime(&ms1, a1, b1);
                CK(cudaSetDevice(0));
            }
            char name[128];
            snprintf(name, sizeof(name), "bidirectional TMA bulk store %5d B pieces: dev0 / dev1", piece);
            report(name, ms0); report(name, ms1);
        }
    }
    return 0;
}
