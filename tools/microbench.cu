// microbench.cu -- measured non-HBM peaks of the B200 the roofline fractions in DESIGN.md / bench.py are quoted against
// ("verify the 16/clk figure with a microbenchmark before quoting").
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench tools/microbench.cu      (tools/build_microbench.sh)
//   tools/_build/microbench > gpurun_out/microbench.json
//
// Every kernel runs ITER dependent-free instruction streams per thread (8 independent chains, so the pipe and not the
// latency is measured), 148 x 8 CTAs of 256 threads, timed with CUDA events after a warm-up launch; the result is
// operations per second over the whole chip and per clock per SM at the SM clock nvidia-smi reports under load.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

#define ITER 4096
#define CHAINS 8

__global__ void k_popc32(uint32_t* out, uint32_t seed) {
    uint32_t v[CHAINS], acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { v[c] = seed + threadIdx.x * 2654435761u + c; acc[c] = 0; }
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) { acc[c] += __popc(v[c] ^ acc[c]); }     // POPC + LOP3 + IADD per chain step
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    if (s == 0xdeadbeef) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the match kernel's inner step: xor + popc.b64 + add on 64-bit words (one 256-bit distance = 4 of these)
__global__ void k_popc64(uint32_t* out, unsigned long long seed) {
    unsigned long long v[CHAINS];
    uint32_t acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { v[c] = seed + threadIdx.x * 0x9E3779B97F4A7C15ull + c; acc[c] = 0; }
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) { acc[c] += __popcll(v[c] ^ ((unsigned long long)acc[c] << 7)); }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    if (s == 0xdeadbeef) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lop3(uint32_t* out, uint32_t seed) {          // alu pipe: LOP3 / IADD3
    uint32_t a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x + c;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) a[c] = (a[c] ^ 0x5bd1e995u) + (a[c] >> 3);   // LOP3/SHF + IADD3: ~3 alu ops
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= a[c];
    if (s == 0xdeadbeef) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad(uint32_t* out, uint32_t seed) {          // fma pipe: IMAD
    uint32_t a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x + c;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) a[c] = a[c] * 1664525u + seed;
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= a[c];
    if (s == 0xdeadbeef) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_vmnmx(uint32_t* out, uint32_t seed) {         // byte-SIMD min / max (the FAST arc test): __vminu4 / __vmaxu4
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { a[c] = seed + threadIdx.x * 0x01010101u + c; b[c] = ~a[c]; }
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) { a[c] = __vminu4(a[c], b[c]); b[c] = __vmaxu4(b[c], a[c] + 0x01010101u); }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= a[c] ^ b[c];
    if (s == 0xdeadbeef) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, double seed) {              // FP64 FMA pipe (bundle adjustment)
    double a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x * 1e-3 + c;
    const double m = 1.0000001, k = 1e-9;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) a[c] = fma(a[c], m, k);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += a[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, float seed) {
    float a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) a[c] = seed + threadIdx.x * 1e-3f + c;
    const float m = 1.0000001f, k = 1e-9f;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) a[c] = fmaf(a[c], m, k);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += a[c];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 tensor-core path: mma.sync m8n8k4 (DMMA.884), 256 FMA per warp instruction, 4 independent accumulator pairs per warp
__global__ void k_dmma(double* out, double seed) {
    double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double a = seed + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[2 * k]), "+d"(c[2 * k + 1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// shared-memory byte loads (FAST ring gathers, rBRIEF sampling): LDS.U8, conflict-free
__global__ void k_lds8(uint32_t* out, int stride) {
    __shared__ uint8_t sm[16384];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = (uint8_t)(i * 7);
    __syncthreads();
    uint32_t acc[CHAINS], idx = threadIdx.x * stride;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = c;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) acc[c] += sm[(idx + c * 1031 + i * 32 + (acc[c] & 1)) & 16383];   // 8 independent chains
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    if (s == 0xdeadbeef) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// L2-resident read bandwidth: every CTA streams a 32 MB window that fits the 126 MB L2 (128-bit loads)
__global__ void k_l2read(const uint4* __restrict__ buf, size_t n16, uint32_t* out, int reps) {
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0xdeadbeef) out[0] = acc;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a)); CHECK(cudaEventCreate(&b));
    launch();                                  // warm-up
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CHECK(cudaEventRecord(a));
        launch();
        CHECK(cudaEventRecord(b));
        CHECK(cudaEventSynchronize(b));
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CHECK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, grid = sms * 8, threads = 256;
    int clk_khz = 0;
    CHECK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double clk = clk_khz * 1e3;          // max SM clock; the JSON also carries ops/s so that any clock can be applied
    void* out;
    CHECK(cudaMalloc(&out, (size_t)grid * threads * 8));
    const double nthread_ops = (double)grid * threads * ITER * CHAINS;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_hz_max\": %.0f", prop.name, sms, clk);
    auto report = [&](const char* name, double ops, double ms, const char* unit) {
        const double per_s = ops / (ms * 1e-3);
        printf(", \"%s\": {\"per_s\": %.4g, \"per_clk_per_sm_at_max_clock\": %.2f, \"ms\": %.4f, \"unit\": \"%s\"}", name, per_s, per_s / clk / sms, ms, unit);
    };
    report("popc_b32", nthread_ops, time_ms([&] { k_popc32<<<grid, threads>>>((uint32_t*)out, 1u); }), "popc.b32 (+xor, +add) thread-ops");
    report("popc_b64", nthread_ops, time_ms([&] { k_popc64<<<grid, threads>>>((uint32_t*)out, 1ull); }), "popc.b64 (+xor, +add) thread-ops; one 256-bit Hamming distance = 4");
    report("alu_lop3_iadd", nthread_ops * 3, time_ms([&] { k_lop3<<<grid, threads>>>((uint32_t*)out, 1u); }), "alu-pipe thread-ops (LOP3 / SHF / IADD3)");
    report("imad", nthread_ops, time_ms([&] { k_imad<<<grid, threads>>>((uint32_t*)out, 3u); }), "IMAD thread-ops");
    report("vminmax_u8x4", nthread_ops * 2, time_ms([&] { k_vmnmx<<<grid, threads>>>((uint32_t*)out, 3u); }), "__vminu4 / __vmaxu4 thread-ops (4 bytes each)");
    report("dfma", nthread_ops, time_ms([&] { k_dfma<<<grid, threads>>>((double*)out, 1.0); }), "FP64 FMA thread-ops (x2 = flop)");
    report("dmma_m8n8k4_fma", (double)grid * (threads / 32) * ITER * 4 * 256, time_ms([&] { k_dmma<<<grid, threads>>>((double*)out, 1.0); }), "FP64 FMA done by mma.sync.m8n8k4 (256 per warp instruction; x2 = flop)");
    report("ffma", nthread_ops, time_ms([&] { k_ffma<<<grid, threads>>>((float*)out, 1.0f); }), "FP32 FMA thread-ops (x2 = flop)");
    report("lds_u8", nthread_ops, time_ms([&] { k_lds8<<<grid, threads>>>((uint32_t*)out, 1); }), "LDS.U8 thread-loads (conflict-free)");
    {
        const size_t bytes = 32u << 20;
        uint4* buf;
        CHECK(cudaMalloc(&buf, bytes));
        CHECK(cudaMemset(buf, 1, bytes));
        const int reps = 20;
        const double ms = time_ms([&] { k_l2read<<<sms * 8, 256>>>(buf, bytes / 16, (uint32_t*)out, reps); });
        printf(", \"l2_read\": {\"GBps\": %.1f, \"window_MB\": 32, \"ms\": %.4f}", (double)bytes * reps / (ms * 1e-3) / 1e9, ms);
        CHECK(cudaFree(buf));
    }
    printf("}\n");
    return 0;
}
