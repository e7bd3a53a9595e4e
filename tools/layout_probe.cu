// Layout probe for the v1 narrow phase (circle/circle pair tests + contact set-up), the phase where "bodies on
// lanes" has the most to offer: one THREAD per environment (the shipped layout: serial loop over the B(B-1)/2
// pairs, positions in shared-memory columns) against one WARP per environment (pairs spread over the 32 lanes,
// positions fetched with warp shuffles, contacts counted with a ballot).  Same inputs, same arithmetic per pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/layout_probe tools/layout_probe.cu && /tmp/layout_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__constant__ unsigned char c_pair_i[256], c_pair_j[256];

__device__ __forceinline__ double pair_work(double ax, double ay, double bx, double by, double mind, int &touch)
{   // what space_step does per pair: distance test; on contact the normal and the penetration
    const double dx = __dsub_rn(bx, ax), dy = __dsub_rn(by, ay);
    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    touch = d2 < mind * mind;
    double pen = 0.0;
    if (touch) {
        const double dist = __dsqrt_rn(d2);
        const double inv = __ddiv_rn(1.0, dist);
        pen = __dadd_rn(__dmul_rn(__dmul_rn(dx, inv), dx), __dmul_rn(__dmul_rn(dy, inv), dy)) - mind;
    }
    return pen;
}

// (A) one thread per environment; positions in shared memory, column per lane (stride 33)
template <int B>
__global__ void thread_per_env(const double *__restrict__ pos, int n, int reps, double *out_pen, int *out_cnt)
{
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *col = sm + warp * (2 * B * 33) + lane;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    for (int k = 0; k < 2 * B; ++k) col[k * 33] = pos[(size_t)k * n + e];
    double pen = 0.0;
    int cnt = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int j = 1; j < B; ++j) {
            const double jx = col[(2 * j) * 33], jy = col[(2 * j + 1) * 33];
#pragma unroll 1
            for (int i = 0; i < j; ++i) {
                int t;
                pen += pair_work(col[(2 * i) * 33], col[(2 * i + 1) * 33], jx, jy, (i == B - 1 || j == B - 1) ? 2.5 : 3.0, t);
                cnt += t;
            }
        }
        col[0] = col[0] + 1e-3;       // the bodies move between steps
    }
    out_pen[e] = pen; out_cnt[e] = cnt;
}

// (B) one warp per environment; lane b holds body b, pair q = 32 * round + lane
template <int B>
__global__ void warp_per_env(const double *__restrict__ pos, int n, int reps, double *out_pen, int *out_cnt)
{
    constexpr int P = B * (B - 1) / 2, ROUNDS = (P + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= n) return;
    double x = lane < B ? pos[(size_t)(2 * lane) * n + e] : 0.0, y = lane < B ? pos[(size_t)(2 * lane + 1) * n + e] : 0.0;
    double pen = 0.0;
    int cnt = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int q = rd * 32 + lane;
            const int i = q < P ? c_pair_i[q] : 0, j = q < P ? c_pair_j[q] : 0;
            const double ax = __shfl_sync(0xffffffffu, x, i), ay = __shfl_sync(0xffffffffu, y, i);
            const double bx = __shfl_sync(0xffffffffu, x, j), by = __shfl_sync(0xffffffffu, y, j);
            int t = 0;
            double p = 0.0;
            if (q < P) p = pair_work(ax, ay, bx, by, (i == B - 1 || j == B - 1) ? 2.5 : 3.0, t);
            pen += p;
            cnt += __popc(__ballot_sync(0xffffffffu, t));      // the contact list is built from the ballot
        }
        if (lane == 0) x = x + 1e-3;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pen += __shfl_xor_sync(0xffffffffu, pen, o);
    if (lane == 0) { out_pen[e] = pen; out_cnt[e] = cnt; }
}

template <int B>
void run(int n, int reps)
{
    std::vector<double> h((size_t)2 * B * n);
    srand(1);
    for (int e = 0; e < n; ++e)
        for (int b = 0; b < B; ++b) {      // players bunch around the ball: ~0.7 contacts per env at 5v5
            h[(size_t)(2 * b) * n + e] = 52.5 + (rand() / (double)RAND_MAX - 0.5) * (b % 3 == 0 ? 12.0 : 90.0);
            h[(size_t)(2 * b + 1) * n + e] = 34.0 + (rand() / (double)RAND_MAX - 0.5) * (b % 3 == 0 ? 12.0 : 60.0);
        }
    unsigned char pi[256], pj[256];
    int q = 0;
    for (int j = 1; j < B; ++j) for (int i = 0; i < j; ++i) { pi[q] = i; pj[q] = j; ++q; }
    cudaMemcpyToSymbol(c_pair_i, pi, q); cudaMemcpyToSymbol(c_pair_j, pj, q);
    double *d_pos, *d_pen; int *d_cnt;
    cudaMalloc(&d_pos, h.size() * 8); cudaMalloc(&d_pen, (size_t)n * 8); cudaMalloc(&d_cnt, (size_t)n * 4);
    cudaMemcpy(d_pos, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms; long long cnt_a = 0, cnt_b = 0; double pen_a = 0, pen_b = 0;
    std::vector<int> hc(n); std::vector<double> hp(n);
    const int tpb = 64, smem = (tpb / 32) * 2 * B * 33 * 8;
    for (int it = 0; it < 2; ++it) {
        cudaEventRecord(a);
        thread_per_env<B><<<(n + tpb - 1) / tpb, tpb, smem>>>(d_pos, n, reps, d_pen, d_cnt);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    }
    cudaMemcpy(hc.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hp.data(), d_pen, (size_t)n * 8, cudaMemcpyDeviceToHost);
    for (int e = 0; e < n; ++e) { cnt_a += hc[e]; pen_a += hp[e]; }
    const double ta = ms;
    for (int it = 0; it < 2; ++it) {
        cudaEventRecord(a);
        warp_per_env<B><<<(int)(((size_t)n * 32 + 255) / 256), 256>>>(d_pos, n, reps, d_pen, d_cnt);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    }
    cudaMemcpy(hc.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hp.data(), d_pen, (size_t)n * 8, cudaMemcpyDeviceToHost);
    for (int e = 0; e < n; ++e) { cnt_b += hc[e]; pen_b += hp[e]; }
    printf("B=%2d bodies (%3d pairs), %d envs x %d steps: thread-per-env %.3f ms (%.3e env-steps/s) | warp-per-env %.3f ms (%.3e env-steps/s) | "
           "contacts/env-step %.3f vs %.3f, checksum %.6f vs %.6f | %s\n", B, B * (B - 1) / 2, n, reps, ta, (double)n * reps / ta * 1e3, ms,
           (double)n * reps / ms * 1e3, (double)cnt_a / n / reps, (double)cnt_b / n / reps, pen_a, pen_b, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_pos); cudaFree(d_pen); cudaFree(d_cnt);
}

int main()
{
    run<5>(1 << 18, 64);
    run<11>(1 << 18, 64);
    run<21>(1 << 16, 64);
    return 0;
}
