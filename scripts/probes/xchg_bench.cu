// Probe (not part of the product): what does it cost, on this GPU, to move a few self-validating
// 8-byte words between CTAs through L2?
//   mode 0  ping-pong between CTA 0 and CTA k: one-way store -> visible-to-a-polling-load latency
//   mode 1  all-to-all: every CTA pushes W words into a private mailbox of every CTA, then polls its
//           own mailbox until all G x W words of the round are there (the exchange of k_resident)
//   mode 2  all-to-all through ONE shared mailbox that every CTA polls (the v1 scheme)
//   mode 3  like 1, but only warp 0 polls, with 128-bit loads
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_bench xchg_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__device__ __forceinline__ void st_word(unsigned long long * p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long * p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_word2(const unsigned long long * p, unsigned long long & a, unsigned long long & b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// mode 0: CTA 0 and CTA `peer` bounce a counter `rounds` times. cycles[0] = total cycles seen by CTA 0.
__global__ void k_pingpong(unsigned long long * box, int peer, int rounds, long long * cycles)
{
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    if (me != 0 && me != peer) return;
    unsigned long long * mine = box + (me == 0 ? 0 : 32);    // 256 B apart
    unsigned long long * theirs = box + (me == 0 ? 32 : 0);
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r)
    {
        if (me == 0)
        {
            st_word(theirs, (unsigned long long)r);
            while (ld_word(mine) < (unsigned long long)r) {}
        }
        else
        {
            while (ld_word(mine) < (unsigned long long)r) {}
            st_word(theirs, (unsigned long long)r);
        }
    }
    if (me == 0) cycles[0] = clock64() - t0;
}

// modes 1-3: `rounds` exchanges back to back; cycles[cta] = cycles per CTA for all rounds
template <int MODE>
__global__ void k_alltoall(unsigned long long * mail, int W, int rounds, long long * cycles, unsigned int seq0)
{
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    __shared__ unsigned int words[4 * 192];
    __shared__ int done;
    cooperative_groups::this_grid().sync();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r)
    {
        const unsigned int seq = seq0 + r;
        const unsigned long long word = ((unsigned long long)seq << 32) | (unsigned)(cta * 16 + (tid & 3));
        if (MODE == 2)
        {
            if (tid < W) st_word(mail + (size_t)cta * 4 + tid, word);
        }
        else
        {
            const int k = tid & 3;
            if (k < W)
                for (int t = tid; t < 4 * G; t += blockDim.x) st_word(mail + ((size_t)(t >> 2) * G + cta) * 4 + k, word);
        }
        const unsigned long long * box = (MODE == 2) ? mail : mail + (size_t)cta * G * 4;
        if (MODE == 3)
        {
            if (tid == 0) done = 0;
            __syncthreads();
            if (tid < 32)
            {
                // 2 slots (64 B) per lane and trip: lanes cover 64 slots per trip
                for (;;)
                {
                    bool ok = true;
                    for (int s = 2 * tid; s < G; s += 64)
                    {
                        unsigned long long a, b, c, d, e, f, g, h;
                        ld_word2(box + (size_t)s * 4, a, b);
                        ld_word2(box + (size_t)s * 4 + 2, c, d);
                        const bool two = s + 1 < G;
                        if (two)
                        {
                            ld_word2(box + (size_t)s * 4 + 4, e, f);
                            ld_word2(box + (size_t)s * 4 + 6, g, h);
                        }
                        ok = ok && (unsigned)(a >> 32) >= seq && (W < 2 || (unsigned)(b >> 32) >= seq) &&
                             (W < 3 || (unsigned)(c >> 32) >= seq) && (W < 4 || (unsigned)(d >> 32) >= seq);
                        if (two)
                            ok = ok && (unsigned)(e >> 32) >= seq && (W < 2 || (unsigned)(f >> 32) >= seq) &&
                                 (W < 3 || (unsigned)(g >> 32) >= seq) && (W < 4 || (unsigned)(h >> 32) >= seq);
                    }
                    if (__all_sync(0xffffffffu, ok)) break;
                }
            }
            __syncthreads();
        }
        else
        {
            const int total = W * G;
            for (int w = tid; w < total; w += blockDim.x)
            {
                const int a = (w / W) * 4 + (w % W);
                unsigned long long v;
                do { v = ld_word(box + a); } while ((unsigned)(v >> 32) < seq);  // >=: a fast peer may already have posted the next round
                words[a] = (unsigned)v;
            }
            __syncthreads();
        }
    }
    if (tid == 0) cycles[cta] = clock64() - t0;
}

int main(int argc, char ** argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int G = prop.multiProcessorCount;
    const double mhz = prop.clockRate / 1000.0;
    std::printf("device %s, %d SMs, %.0f MHz\n", prop.name, G, mhz);
    unsigned long long * mail;
    long long * cycles;
    const size_t words = (size_t)G * G * 4 + 4096;
    CK(cudaMalloc(&mail, words * 8));
    CK(cudaMalloc(&cycles, G * sizeof(long long)));
    std::vector<long long> h(G);

    // mode 0
    {
        const int rounds = 2000;
        std::vector<double> lat;
        for (int peer = 1; peer < G; ++peer)
        {
            CK(cudaMemset(mail, 0, 4096));
            k_pingpong<<<G, 32>>>(mail, peer, rounds, cycles);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), cycles, sizeof(long long), cudaMemcpyDeviceToHost));
            lat.push_back(h[0] / (2.0 * rounds) / mhz);
        }
        std::sort(lat.begin(), lat.end());
        std::printf("ping-pong one-way store->seen-by-poll latency over %d peers: min %.3f  median %.3f  max %.3f us\n", G - 1,
                    lat.front(), lat[lat.size() / 2], lat.back());
    }
    // modes 1-3
    unsigned int seq = 1;
    for (int mode = 1; mode <= 3; ++mode)
        for (int W = 1; W <= 4; W += 3)
            for (int threads = 128; threads <= 512; threads *= 2)
            {
                if (mode == 3 && threads != 512) continue;
                const int rounds = 2000;
                CK(cudaMemset(mail, 0, words * 8));
                int Wv = W, rv = rounds;
                unsigned int sv = seq;
                void * args[] = { &mail, &Wv, &rv, &cycles, &sv };
                const void * fn = mode == 1 ? (const void *)k_alltoall<1> : mode == 2 ? (const void *)k_alltoall<2> : (const void *)k_alltoall<3>;
                CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(threads), args, 0, 0));
                CK(cudaDeviceSynchronize());
                seq += rounds;
                CK(cudaMemcpy(h.data(), cycles, G * sizeof(long long), cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (int c = 0; c < G; ++c) mx = std::max(mx, h[c]);
                std::printf("mode %d (%s)  W=%d words  %3d threads: %.3f us per exchange\n", mode,
                            mode == 1 ? "private mailboxes, all threads poll" : mode == 2 ? "one shared mailbox" : "private mailboxes, warp 0 polls 128-bit",
                            W, threads, mx / (double)rounds / mhz);
            }
    return 0;
}
