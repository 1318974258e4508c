// orb_bow.cu -- sm_100a DBoW2 image -> (BowVector, FeatureVector) conversion behind orbv_create / orbv_transform.
//
// Reference path (file:line under /root/reference): Frame::ComputeBoW src/Frame.cc:393-408 (transform(vCurrentDesc, mvBowVec[c],
// mvFeatVec[c], 4) per camera); TemplatedVocabulary::transform Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1149-1227 (image) and
// :1249-1292 (one feature: tree descent, k Hamming distances per level, first child wins ties); loadFromTextFile :1362-1447;
// FORB::distance Thirdparty/DBoW2/DBoW2/FORB.cpp:82-102; BowVector::addWeight / normalize(L1) BowVector.cpp:36-88.
//
// Device formulation, batched over descriptor sets (one set = one camera image):
//   k_descend   one thread per feature: walks the tree; the children of a node are stored contiguously (descriptor rows re-ordered at
//               create), so a level is k consecutive 32-byte rows -- the upper levels stay in L1 / L2, the leaf level of the ORB
//               vocabulary (10^6 rows = 32 MB) fits the 126 MB L2.
//   k_sets      one CTA per set: bitonic sort of (word id, feature) and (node id, feature) keys in shared memory = the iteration
//               order of the two std::maps; run-length heads give the BowVector entries (value = weight added once per feature, in
//               feature order, as addWeight does) and the FeatureVector CSR; the L1 norm is summed by ONE thread in ascending word
//               order, because the reference's double sum is order dependent and the result is compared bit for bit.
// There is no CPU fallback.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "orb_common.h"

#define BOW_MAX_SET 8192          // features per descriptor set (one image); 2 x 64 KB of keys in shared memory
#define BOW_T 1024

struct orbv {
    int device = 0, k = 0, L = 0, n_nodes = 0, n_words = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int* d_child_off = nullptr;    // [n_nodes + 1] into the child arrays
    int* d_child_id = nullptr;     // [n_children] node id of every child, children of a node consecutive, in file order
    uint4* d_child_desc = nullptr; // [n_children][2]
    double* d_weight = nullptr;    // [n_nodes]
    int* d_word = nullptr;         // [n_nodes] word id of a leaf, -1 otherwise
    uint8_t* d_buf = nullptr; size_t cap = 0;
    uint8_t* h_buf = nullptr; size_t hcap = 0;
    long long launches = 0;
};

__device__ __forceinline__ int hamming256(uint4 a0, uint4 a1, uint4 b0, uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) +
           __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(128) k_descend(const uint4* __restrict__ desc, int n, int nid_level, const int* __restrict__ child_off,
                                                 const int* __restrict__ child_id, const uint4* __restrict__ child_desc, const double* __restrict__ weight,
                                                 const int* __restrict__ word, int* __restrict__ word_id, int* __restrict__ node_id, double* __restrict__ w_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 f0 = desc[2 * (size_t)i], f1 = desc[2 * (size_t)i + 1];
    int nid = 0, final_id = 0, level = 0;
    int a = child_off[0], b = child_off[1];
    while (b > a) {                                    // do { ... } while(!isLeaf())
        ++level;
        int best_d = 1 << 30, best = a;
        for (int c = a; c < b; c++) {
            const int d = hamming256(f0, f1, child_desc[2 * (size_t)c], child_desc[2 * (size_t)c + 1]);
            if (d < best_d) { best_d = d; best = c; }  // strict: the first child wins ties
        }
        final_id = child_id[best];
        if (level == nid_level) nid = final_id;
        a = child_off[final_id]; b = child_off[final_id + 1];
    }
    word_id[i] = word[final_id]; node_id[i] = nid; w_out[i] = weight[final_id];
}

__device__ void bitonic_sort(unsigned long long* key, int P) {
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += BOW_T) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = key[i], y = key[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { key[i] = y; key[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

// exclusive prefix of one int per thread over the CTA
__device__ int block_excl_scan(int v, int* total) {
    __shared__ int wsum[32];
    __shared__ int tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) wsum[w] = s;
    __syncthreads();
    if (w == 0) {
        int x = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
        wsum[lane] = x;
        if (lane == 31) tot = x;
    }
    __syncthreads();
    const int before = (w ? wsum[w - 1] : 0) + s - v;
    *total = tot;
    __syncthreads();
    return before;
}

// one CTA per descriptor set
__global__ void __launch_bounds__(BOW_T) k_sets(const int* __restrict__ set_off, const int* __restrict__ word_id, const int* __restrict__ node_id,
                                                const double* __restrict__ w, int* __restrict__ bow_ids, double* __restrict__ bow_vals,
                                                int* __restrict__ n_words, int* __restrict__ fv_node, int* __restrict__ fv_off, int* __restrict__ fv_idx,
                                                int* __restrict__ n_fv_nodes) {
    extern __shared__ unsigned long long key[];
    const int s = blockIdx.x, lo = set_off[s], n = set_off[s + 1] - lo;
    int P = 1;
    while (P < n) P <<= 1;
    const int per = (P + BOW_T - 1) / BOW_T, c0 = threadIdx.x * per;
    for (int pass = 0; pass < 2; pass++) {
        const int* id = pass == 0 ? word_id : node_id;
        for (int i = threadIdx.x; i < P; i += BOW_T)
            key[i] = (i < n && w[lo + i] > 0.0) ? (((unsigned long long)(unsigned)id[lo + i] << 32) | (unsigned)i) : ~0ull;       // `if(w > 0)`: not stopped
        __syncthreads();
        bitonic_sort(key, P);
        // heads of the runs of equal ids inside this thread's chunk
        int heads = 0;
        for (int i = c0; i < c0 + per && i < P; i++) {
            if (key[i] == ~0ull) break;
            heads += (i == 0 || (key[i] >> 32) != (key[i - 1] >> 32));
        }
        int total;
        int pos = block_excl_scan(heads, &total);
        if (pass == 0) {
            for (int i = c0; i < c0 + per && i < P; i++) {
                if (key[i] == ~0ull) break;
                if (i == 0 || (key[i] >> 32) != (key[i - 1] >> 32)) {
                    const double wt = w[lo + (int)(key[i] & 0xffffffffu)];
                    double v = wt;                                           // insert(id, w), then `vit->second += v` once per further feature
                    for (int j = i + 1; j < P && (key[j] >> 32) == (key[i] >> 32); j++) v = __dadd_rn(v, wt);
                    bow_ids[lo + pos] = (int)(key[i] >> 32); bow_vals[lo + pos] = v;
                    pos++;
                }
            }
            __syncthreads();
            __shared__ double s_norm;
            if (threadIdx.x == 0) {                                          // BowVector::normalize(L1): ascending id order
                double norm = 0.0;
                for (int j = 0; j < total; j++) norm = __dadd_rn(norm, fabs(bow_vals[lo + j]));
                s_norm = norm;
                n_words[s] = total;
            }
            __syncthreads();
            if (s_norm > 0.0) for (int j = threadIdx.x; j < total; j += BOW_T) bow_vals[lo + j] = __ddiv_rn(bow_vals[lo + j], s_norm);
        } else {
            int* off = fv_off + lo + s;                                      // n + 1 entries per set
            int valid = 0;
            for (int i = c0; i < c0 + per && i < P; i++) {
                if (key[i] == ~0ull) break;
                valid++;
                fv_idx[lo + i] = (int)(key[i] & 0xffffffffu);
                if (i == 0 || (key[i] >> 32) != (key[i - 1] >> 32)) { fv_node[lo + pos] = (int)(key[i] >> 32); off[pos] = i; pos++; }
            }
            int nvalid;
            block_excl_scan(valid, &nvalid);
            if (threadIdx.x == 0) { off[total] = nvalid; n_fv_nodes[s] = total; }
        }
        __syncthreads();
    }
}

static void orbv_free(orbv* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d_child_off); cudaFree(v->d_child_id); cudaFree(v->d_child_desc); cudaFree(v->d_weight); cudaFree(v->d_word); cudaFree(v->d_buf);
    if (v->h_buf) cudaFreeHost(v->h_buf);
    if (v->own_stream) cudaStreamDestroy(v->own_stream);
    delete v;
}

extern "C" {

int orbv_create(orbv_t** out, int device, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
    if (!out) ORB_FAIL(ORB_E_INVALID, "orbv_create: out is NULL");
    *out = nullptr;
    if (k < 0 || k > 20 || L < 1 || L > 10 || n_nodes < 2 || !parent || !is_leaf || !desc || !weight)                // the loader's own limits  :1383
        ORB_FAIL(ORB_E_INVALID, "orbv_create: bad vocabulary (k in 0..20, L in 1..10, at least one node below the root)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) ORB_FAIL(ORB_E_NO_DEVICE, "orbv_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) ORB_FAIL(ORB_E_NO_DEVICE, "orbv_create: device %d not present", device);
    cudaDeviceProp prop;
    ORB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) ORB_FAIL(ORB_E_NO_DEVICE, "orbv_create: device %d is sm_%d%d, the kernels are built for sm_100a only", device, prop.major, prop.minor);
    // children lists in file order (m_nodes[pid].children.push_back(nid)); word ids in file order of the leaves
    std::vector<int> cnt((size_t)n_nodes + 1, 0), word((size_t)n_nodes, -1);
    int n_words = 0;
    for (int nid = 1; nid < n_nodes; nid++) {
        if (parent[nid] < 0 || parent[nid] >= nid) ORB_FAIL(ORB_E_INVALID, "orbv_create: node %d names parent %d (a parent precedes its children in the file)", nid, parent[nid]);
        cnt[parent[nid] + 1]++;
        if (is_leaf[nid]) word[nid] = n_words++;
    }
    for (int i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
    for (int nid = 1; nid < n_nodes; nid++)
        if (is_leaf[nid] != (cnt[nid + 1] == cnt[nid])) ORB_FAIL(ORB_E_INVALID, "orbv_create: node %d: leaf flag and children disagree", nid);
    if (cnt[1] == cnt[0]) ORB_FAIL(ORB_E_INVALID, "orbv_create: the root has no children");
    const int n_children = n_nodes - 1;
    std::vector<int> child_id((size_t)n_children), fill(cnt.begin(), cnt.end() - 1);
    std::vector<uint8_t> cdesc(32 * (size_t)n_children);
    for (int nid = 1; nid < n_nodes; nid++) {
        const int slot = fill[parent[nid]]++;
        child_id[slot] = nid;
        memcpy(cdesc.data() + 32 * (size_t)slot, desc + 32 * (size_t)nid, 32);
    }
    ORB_CUDA(cudaSetDevice(device));
    orbv* v = new (std::nothrow) orbv();
    if (!v) ORB_FAIL(ORB_E_INVALID, "orbv_create: out of host memory");
    v->device = device; v->k = k; v->L = L; v->n_nodes = n_nodes; v->n_words = n_words;
    cudaError_t ce = cudaStreamCreateWithFlags(&v->own_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v->d_child_off, 4 * ((size_t)n_nodes + 1));
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v->d_child_id, 4 * (size_t)n_children);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v->d_child_desc, 32 * (size_t)n_children);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v->d_weight, 8 * (size_t)n_nodes);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&v->d_word, 4 * (size_t)n_nodes);
    if (ce == cudaSuccess) ce = cudaMemcpy(v->d_child_off, cnt.data(), 4 * ((size_t)n_nodes + 1), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(v->d_child_id, child_id.data(), 4 * (size_t)n_children, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(v->d_child_desc, cdesc.data(), 32 * (size_t)n_children, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(v->d_weight, weight, 8 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(v->d_word, word.data(), 4 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_sets, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * BOW_MAX_SET);
    if (ce != cudaSuccess) { const int rc = orbhost::check_cuda(ce, "orbv_create", __FILE__, __LINE__); orbv_free(v); return rc; }
    v->stream = v->own_stream;
    *out = v;
    return ORB_OK;
}

void orbv_destroy(orbv_t* v) { orbv_free(v); }
int orbv_words(const orbv_t* v) { return v ? v->n_words : 0; }
long long orbv_launch_count(const orbv_t* v) { return v ? v->launches : 0; }
int orbv_set_stream(orbv_t* v, void* cuda_stream) {
    if (!v) ORB_FAIL(ORB_E_INVALID, "orbv_set_stream: NULL handle");
    v->stream = cuda_stream ? (cudaStream_t)cuda_stream : v->own_stream;
    return ORB_OK;
}

int orbv_transform(orbv_t* v, const uint8_t* desc, const int32_t* set_off, int n_sets, int levelsup, int32_t* word_id, int32_t* node_id, int32_t* bow_ids,
                   double* bow_vals, int32_t* n_words, int32_t* fv_node, int32_t* fv_off, int32_t* fv_idx, int32_t* n_fv_nodes) {
    if (!v) ORB_FAIL(ORB_E_INVALID, "orbv_transform: NULL handle");
    if (n_sets < 0 || !set_off || !bow_ids || !bow_vals || !n_words || !fv_node || !fv_off || !fv_idx || !n_fv_nodes) ORB_FAIL(ORB_E_INVALID, "orbv_transform: bad argument");
    if (n_sets == 0) return ORB_OK;
    if (set_off[0] != 0) ORB_FAIL(ORB_E_INVALID, "orbv_transform: set_off[0] must be 0");
    for (int s = 0; s < n_sets; s++) {
        const int m = set_off[s + 1] - set_off[s];
        if (m < 0 || m > BOW_MAX_SET) ORB_FAIL(ORB_E_INVALID, "orbv_transform: set %d has %d features (0..%d)", s, m, BOW_MAX_SET);
    }
    const int n = set_off[n_sets];
    if (n && !desc) ORB_FAIL(ORB_E_INVALID, "orbv_transform: NULL descriptors");
    ORB_CUDA(cudaSetDevice(v->device));
    cudaStream_t st = v->stream;
    // layout (device == pinned mirror): desc | set_off || word | node | w | bow_ids | bow_vals | n_words | fv_node | fv_off | fv_idx | n_fv
    size_t cur = 0;
    auto add = [&](size_t b) { const size_t o = (cur + 255) & ~(size_t)255; cur = o + b; return o; };
    const size_t o_desc = add(32 * (size_t)n), o_soff = add(4 * ((size_t)n_sets + 1));
    const size_t staged = add(0);
    const size_t o_word = add(4 * (size_t)n), o_node = add(4 * (size_t)n), o_w = add(8 * (size_t)n), o_bid = add(4 * (size_t)n), o_bval = add(8 * (size_t)n),
                 o_nw = add(4 * (size_t)n_sets), o_fvn = add(4 * (size_t)n), o_fvo = add(4 * ((size_t)n + n_sets)), o_fvi = add(4 * (size_t)n), o_nf = add(4 * (size_t)n_sets);
    const size_t total = add(0);
    if (total > v->cap) {
        if (v->d_buf) cudaFree(v->d_buf);
        v->d_buf = nullptr; v->cap = 0;
        ORB_CUDA(cudaMalloc((void**)&v->d_buf, total + total / 2));
        v->cap = total + total / 2;
    }
    if (total > v->hcap) {
        if (v->h_buf) cudaFreeHost(v->h_buf);
        v->h_buf = nullptr; v->hcap = 0;
        ORB_CUDA(cudaHostAlloc((void**)&v->h_buf, total + total / 2, cudaHostAllocDefault));
        v->hcap = total + total / 2;
    }
    uint8_t *D = v->d_buf, *H = v->h_buf;
    if (n) memcpy(H + o_desc, desc, 32 * (size_t)n);
    memcpy(H + o_soff, set_off, 4 * ((size_t)n_sets + 1));
    ORB_CUDA(cudaMemcpyAsync(D, H, staged, cudaMemcpyHostToDevice, st));
    if (n) k_descend<<<(n + 127) / 128, 128, 0, st>>>((const uint4*)(D + o_desc), n, v->L - levelsup, v->d_child_off, v->d_child_id, v->d_child_desc, v->d_weight,
                                                      v->d_word, (int*)(D + o_word), (int*)(D + o_node), (double*)(D + o_w));
    k_sets<<<n_sets, BOW_T, 8 * BOW_MAX_SET, st>>>((const int*)(D + o_soff), (const int*)(D + o_word), (const int*)(D + o_node), (const double*)(D + o_w),
                                                   (int*)(D + o_bid), (double*)(D + o_bval), (int*)(D + o_nw), (int*)(D + o_fvn), (int*)(D + o_fvo),
                                                   (int*)(D + o_fvi), (int*)(D + o_nf));
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(H + o_word, D + o_word, total - o_word, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    v->launches += n ? 2 : 1;
    if (word_id && n) memcpy(word_id, H + o_word, 4 * (size_t)n);
    if (node_id && n) memcpy(node_id, H + o_node, 4 * (size_t)n);
    memcpy(n_words, H + o_nw, 4 * (size_t)n_sets);
    memcpy(n_fv_nodes, H + o_nf, 4 * (size_t)n_sets);
    memcpy(fv_off, H + o_fvo, 4 * ((size_t)n + n_sets));
    // entries past a set's counts are unspecified on the device; copy the defined prefixes only
    for (int s = 0; s < n_sets; s++) {
        const int lo = set_off[s], nw = n_words[s], nf = n_fv_nodes[s];
        memcpy(bow_ids + lo, H + o_bid + 4 * (size_t)lo, 4 * (size_t)nw);
        memcpy(bow_vals + lo, H + o_bval + 8 * (size_t)lo, 8 * (size_t)nw);
        memcpy(fv_node + lo, H + o_fvn + 4 * (size_t)lo, 4 * (size_t)nf);
        const int nvalid = nf ? fv_off[lo + s + nf] : 0;
        memcpy(fv_idx + lo, H + o_fvi + 4 * (size_t)lo, 4 * (size_t)nvalid);
    }
    return ORB_OK;
}

}  // extern "C"
