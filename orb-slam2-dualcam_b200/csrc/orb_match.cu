// orb_match.cu -- sm_100a Hamming matchers behind the orbm_* C-ABI.
//
// Reference path (file:line under /root/reference): ORBmatcher::DescriptorDistance src/ORBmatcher.cc:2015-2031 and the
// best / second-best scans that every SearchBy* runs over it (e.g. src/ORBmatcher.cc:575-610, :216-231).
//
//   hamming_pairs_kernel   DescriptorDistance for n independent descriptor pairs
//   bruteforce_kernel      per query: nearest + second-nearest train descriptor of the same (frame, camera) pair.
//                          One thread owns one query (256 bits in 4 x 64-bit registers); the train set streams through
//                          shared memory in chunks and every thread of the CTA reads the same train descriptor
//                          (broadcast, conflict-free); distance = 4 x popcll; the running (best, second) pair is kept
//                          as packed (distance << 22 | index) keys so that two integer min/max track it exactly.
// There is no CPU fallback.
#include <new>

#include "orb_common.h"

#define BF_THREADS 128
#define BF_CHUNK 512          // train descriptors per shared-memory stage (16 KB)
#define BF_IDX_BITS 22
#define BF_IDX_MASK ((1u << BF_IDX_BITS) - 1u)

__global__ void __launch_bounds__(256) hamming_pairs_kernel(const ulonglong4* __restrict__ a, const ulonglong4* __restrict__ b, int n,
                                                            int* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ulonglong4 x = a[i], y = b[i];
    dist[i] = __popcll(x.x ^ y.x) + __popcll(x.y ^ y.y) + __popcll(x.z ^ y.z) + __popcll(x.w ^ y.w);
}

__global__ void __launch_bounds__(BF_THREADS) bruteforce_kernel(const uint8_t* __restrict__ dq, const int* __restrict__ nq, int q_capacity,
                                                                const uint8_t* __restrict__ dt, const int* __restrict__ nt, int t_capacity,
                                                                int* __restrict__ best_idx, int* __restrict__ best_d, int* __restrict__ second_d,
                                                                const int* __restrict__ q_set, const int* __restrict__ t_set) {
    __shared__ ulonglong2 strain[BF_CHUNK * 2];
    const int pair = blockIdx.y;
    const int qs = q_set ? q_set[pair] : pair, ts = t_set ? t_set[pair] : pair;
    const int NQ = min(nq[qs], q_capacity), NT = min(nt[ts], t_capacity);
    const int q0 = blockIdx.x * BF_THREADS;
    if (q0 >= NQ) return;
    const int q = q0 + threadIdx.x;
    const bool active = q < NQ;
    unsigned long long q0w = 0, q1w = 0, q2w = 0, q3w = 0;
    if (active) {
        const ulonglong2* Q = reinterpret_cast<const ulonglong2*>(dq + ((size_t)qs * q_capacity + q) * 32);
        const ulonglong2 lo = __ldg(Q), hi = __ldg(Q + 1);
        q0w = lo.x; q1w = lo.y; q2w = hi.x; q3w = hi.y;
    }
    const ulonglong2* T = reinterpret_cast<const ulonglong2*>(dt + (size_t)ts * t_capacity * 32);
    unsigned best = (256u << BF_IDX_BITS) | BF_IDX_MASK, second = best;
    for (int t0 = 0; t0 < NT; t0 += BF_CHUNK) {
        const int cn = min(BF_CHUNK, NT - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn * 2; i += BF_THREADS) strain[i] = __ldg(T + (size_t)t0 * 2 + i);
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < cn; t++) {
            const ulonglong2 lo = strain[2 * t], hi = strain[2 * t + 1];
            const unsigned d = __popcll(q0w ^ lo.x) + __popcll(q1w ^ lo.y) + __popcll(q2w ^ hi.x) + __popcll(q3w ^ hi.y);
            const unsigned key = (d << BF_IDX_BITS) | (unsigned)(t0 + t);
            second = min(second, max(best, key));
            best = min(best, key);
        }
    }
    if (active) {
        const size_t o = (size_t)pair * q_capacity + q;
        const int bd = (int)(best >> BF_IDX_BITS);
        best_d[o] = bd;
        best_idx[o] = bd >= 256 ? -1 : (int)(best & BF_IDX_MASK);   // `dist < bestDist` with bestDist = 256 never fires
        second_d[o] = (int)(second >> BF_IDX_BITS);
    }
}

struct orbm {
    int device = 0;
    int max_pairs = 0, max_query = 0, max_train = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    uint8_t *d_q = nullptr, *d_t = nullptr;
    int *d_nq = nullptr, *d_nt = nullptr, *d_out = nullptr;   // d_out: best_idx | best_d | second_d
    long long launches = 0;
    bool profile = false;
    cudaEvent_t prof_ev[2 * 256] = {};
    bool prof_made = false;
    int prof_used = 0;
};

static void orbm_free(orbm* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->prof_made) for (cudaEvent_t ev : m->prof_ev) cudaEventDestroy(ev);
    cudaFree(m->d_q); cudaFree(m->d_t); cudaFree(m->d_nq); cudaFree(m->d_nt); cudaFree(m->d_out);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

// accessors for orb_search.cu (the guided searches share the matcher handle)
cudaStream_t orbm_stream_of(orbm* m) { return m->stream; }
int orbm_device_of(orbm* m) { return m->device; }
void orbm_count_launches(orbm* m, int n) { m->launches += n; }

extern "C" {

int orbm_create(orbm_t** out, int device, int max_pairs, int max_query, int max_train) {
    if (!out) ORB_FAIL(ORB_E_INVALID, "orbm_create: out is NULL");
    *out = nullptr;
    if (max_pairs < 1 || max_pairs > 65535 || max_query < 1 || max_train < 1 || max_train > (int)BF_IDX_MASK)
        ORB_FAIL(ORB_E_INVALID, "orbm_create: sizes out of range (pairs<=65535, train<%u)", BF_IDX_MASK);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) ORB_FAIL(ORB_E_NO_DEVICE, "orbm_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) ORB_FAIL(ORB_E_NO_DEVICE, "orbm_create: device %d not present", device);
    cudaDeviceProp prop;
    ORB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) ORB_FAIL(ORB_E_NO_DEVICE, "orbm_create: device %d is sm_%d%d, the kernels are built for sm_100a only", device, prop.major, prop.minor);
    ORB_CUDA(cudaSetDevice(device));
    orbm* m = new (std::nothrow) orbm();
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_create: out of host memory");
    m->device = device; m->max_pairs = max_pairs; m->max_query = max_query; m->max_train = max_train;
    cudaError_t ce = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { int rc = orbhost::check_cuda(ce, "cudaStreamCreate", __FILE__, __LINE__); orbm_free(m); return rc; }
    m->stream = m->own_stream;
    *out = m;
    return ORB_OK;
}

void orbm_destroy(orbm_t* m) { orbm_free(m); }

int orbm_set_stream(orbm_t* m, void* s) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_set_stream: NULL handle");
    m->stream = s ? (cudaStream_t)s : m->own_stream;
    return ORB_OK;
}

int orbm_synchronize(orbm_t* m) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_synchronize: NULL handle");
    ORB_CUDA(cudaSetDevice(m->device));
    ORB_CUDA(cudaStreamSynchronize(m->stream));
    return ORB_OK;
}

long long orbm_launch_count(const orbm_t* m) { return m ? m->launches : 0; }

static int ensure_staging(orbm* m) {
    if (m->d_q) return ORB_OK;
    const size_t P = (size_t)m->max_pairs;
    ORB_CUDA(cudaMalloc((void**)&m->d_q, P * m->max_query * 32));
    ORB_CUDA(cudaMalloc((void**)&m->d_t, P * m->max_train * 32));
    ORB_CUDA(cudaMalloc((void**)&m->d_nq, P * sizeof(int)));
    ORB_CUDA(cudaMalloc((void**)&m->d_nt, P * sizeof(int)));
    ORB_CUDA(cudaMalloc((void**)&m->d_out, 3 * P * m->max_query * sizeof(int)));
    return ORB_OK;
}

int orbm_descriptor_distance(orbm_t* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_descriptor_distance: NULL handle");
    if (n == 0) return ORB_OK;
    if (!a || !b || !dist || n < 0) ORB_FAIL(ORB_E_INVALID, "orbm_descriptor_distance: bad argument");
    ORB_CUDA(cudaSetDevice(m->device));
    uint8_t *da = nullptr, *db = nullptr;
    int* dd = nullptr;
    int rc = ORB_OK;
    cudaError_t ce = cudaMalloc((void**)&da, (size_t)n * 32);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&db, (size_t)n * 32);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&dd, (size_t)n * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream);
    if (ce == cudaSuccess) {
        hamming_pairs_kernel<<<(n + 255) / 256, 256, 0, m->stream>>>(reinterpret_cast<const ulonglong4*>(da), reinterpret_cast<const ulonglong4*>(db), n, dd);
        m->launches++;
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(dist, dd, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, m->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(m->stream);
    if (ce != cudaSuccess) rc = orbhost::check_cuda(ce, "orbm_descriptor_distance", __FILE__, __LINE__);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return rc;
}

int orbm_bruteforce_device(orbm_t* m, const uint8_t* d_dq, const int32_t* d_nq, int q_capacity, const uint8_t* d_dt, const int32_t* d_nt,
                           int t_capacity, int pairs, int32_t* d_best_idx, int32_t* d_best_d, int32_t* d_second_d) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: NULL handle");
    if (pairs == 0) return ORB_OK;
    if (!d_dq || !d_nq || !d_dt || !d_nt || !d_best_idx || !d_best_d || !d_second_d) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: NULL buffer");
    if (pairs < 0 || pairs > m->max_pairs || q_capacity < 1 || q_capacity > m->max_query || t_capacity < 1 || t_capacity > m->max_train)
        ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: sizes exceed the handle's capacities");
    if (((uintptr_t)d_dq | (uintptr_t)d_dt) & 15) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: descriptor buffers must be 16-byte aligned");
    ORB_CUDA(cudaSetDevice(m->device));
    dim3 grid((q_capacity + BF_THREADS - 1) / BF_THREADS, pairs);
    cudaEvent_t* pev = (m->profile && m->prof_used < 256) ? &m->prof_ev[2 * m->prof_used] : nullptr;
    if (pev) ORB_CUDA(cudaEventRecord(pev[0], m->stream));
    bruteforce_kernel<<<grid, BF_THREADS, 0, m->stream>>>(d_dq, d_nq, q_capacity, d_dt, d_nt, t_capacity, d_best_idx, d_best_d, d_second_d, nullptr, nullptr);
    m->launches++;
    if (pev) { ORB_CUDA(cudaEventRecord(pev[1], m->stream)); m->prof_used++; }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orbm_bruteforce_sets_device(orbm_t* m, const uint8_t* d_desc, const int32_t* d_counts, int capacity, int n_sets,
                                const int32_t* d_q_set, const int32_t* d_t_set, int pairs,
                                int32_t* d_best_idx, int32_t* d_best_d, int32_t* d_second_d) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce_sets: NULL handle");
    if (pairs == 0) return ORB_OK;
    if (!d_desc || !d_counts || !d_q_set || !d_t_set || !d_best_idx || !d_best_d || !d_second_d) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce_sets: NULL buffer");
    if (pairs < 0 || pairs > m->max_pairs || n_sets < 1 || capacity < 1 || capacity > m->max_query || capacity > m->max_train)
        ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce_sets: sizes exceed the handle's capacities");
    if ((uintptr_t)d_desc & 15) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce_sets: descriptor buffer must be 16-byte aligned");
    ORB_CUDA(cudaSetDevice(m->device));
    dim3 grid((capacity + BF_THREADS - 1) / BF_THREADS, pairs);
    cudaEvent_t* pev = (m->profile && m->prof_used < 256) ? &m->prof_ev[2 * m->prof_used] : nullptr;
    if (pev) ORB_CUDA(cudaEventRecord(pev[0], m->stream));
    bruteforce_kernel<<<grid, BF_THREADS, 0, m->stream>>>(d_desc, d_counts, capacity, d_desc, d_counts, capacity, d_best_idx, d_best_d, d_second_d, d_q_set, d_t_set);
    m->launches++;
    if (pev) { ORB_CUDA(cudaEventRecord(pev[1], m->stream)); m->prof_used++; }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orbm_profile(orbm_t* m, int enable) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_profile: NULL handle");
    ORB_CUDA(cudaSetDevice(m->device));
    if (enable && !m->prof_made) {
        for (cudaEvent_t& ev : m->prof_ev) ORB_CUDA(cudaEventCreate(&ev));
        m->prof_made = true;
    }
    m->profile = enable != 0;
    m->prof_used = 0;
    return ORB_OK;
}

int orbm_stage_ms(orbm_t* m, double* ms1, int* calls) {
    if (!m || !ms1) ORB_FAIL(ORB_E_INVALID, "orbm_stage_ms: bad argument");
    ORB_CUDA(cudaSetDevice(m->device));
    ORB_CUDA(cudaStreamSynchronize(m->stream));
    *ms1 = 0.0;
    for (int i = 0; i < m->prof_used; i++) {
        float ms = 0.f;
        ORB_CUDA(cudaEventElapsedTime(&ms, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]));
        *ms1 += ms;
    }
    if (calls) *calls = m->prof_used;
    m->prof_used = 0;
    return ORB_OK;
}

int orbm_bruteforce(orbm_t* m, const uint8_t* dq, const int32_t* nq, int q_capacity, const uint8_t* dt, const int32_t* nt, int t_capacity,
                    int pairs, int32_t* best_idx, int32_t* best_d, int32_t* second_d) {
    if (!m) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: NULL handle");
    if (pairs == 0) return ORB_OK;
    if (!dq || !nq || !dt || !nt || !best_idx || !best_d || !second_d) ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: NULL buffer");
    if (pairs < 0 || pairs > m->max_pairs || q_capacity < 1 || q_capacity > m->max_query || t_capacity < 1 || t_capacity > m->max_train)
        ORB_FAIL(ORB_E_INVALID, "orbm_bruteforce: sizes exceed the handle's capacities");
    ORB_CUDA(cudaSetDevice(m->device));
    int rc = ensure_staging(m);
    if (rc != ORB_OK) return rc;
    const size_t P = (size_t)pairs, nout = P * q_capacity;
    cudaStream_t st = m->stream;
    ORB_CUDA(cudaMemcpyAsync(m->d_q, dq, P * q_capacity * 32, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(m->d_t, dt, P * t_capacity * 32, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(m->d_nq, nq, P * sizeof(int), cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(m->d_nt, nt, P * sizeof(int), cudaMemcpyHostToDevice, st));
    // entries >= nq[p] must be left untouched on the host: seed the device copy with the caller's values
    ORB_CUDA(cudaMemcpyAsync(m->d_out, best_idx, nout * sizeof(int), cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(m->d_out + nout, best_d, nout * sizeof(int), cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(m->d_out + 2 * nout, second_d, nout * sizeof(int), cudaMemcpyHostToDevice, st));
    rc = orbm_bruteforce_device(m, m->d_q, m->d_nq, q_capacity, m->d_t, m->d_nt, t_capacity, pairs, m->d_out, m->d_out + nout, m->d_out + 2 * nout);
    if (rc != ORB_OK) return rc;
    ORB_CUDA(cudaMemcpyAsync(best_idx, m->d_out, nout * sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(best_d, m->d_out + nout, nout * sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(second_d, m->d_out + 2 * nout, nout * sizeof(int), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

}  // extern "C"
