// Host side of libraiko_kzg.so: settings decoding, window-table construction, the
// chunked multi-stream batch scheduler, multi-GPU sharding and the C ABI declared in
// include/raiko_kzg.h.  No CPU arithmetic lives here: every field / curve operation
// runs in the kernels of kzg_kernels.cuh; without a CUDA device the entry points
// fail with RK_ERR_CUDA (no fallback, by design).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/raiko_kzg.h"
#include "kzg_launch.h"

using namespace rk;

namespace {

thread_local std::string g_last_error;

rk_status fail(rk_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return st;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(RK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

// Every extern "C" entry point that touches a device restores the caller's current device on
// exit: a PyTorch (or any other CUDA) host thread must not find its device switched under it.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Per-device scratch arena for DevBuf: one cached allocation that per-call scratch is carved from, so a
// verification call does not pay ~16 cudaMalloc + cudaFree (each cudaFree synchronises the device: 10 ms of a
// 47 ms call).  Only touched under the device mutex.  Grows to the largest call seen.
struct Arena {
    char* base = nullptr;
    size_t cap = 0;
    bool busy = false;
};

struct DevBuf {          // RAII for device scratch that lives for one call: released on every exit path
    std::vector<void*> ptrs;     // individually allocated (no arena, arena busy, or arena too small)
    Arena* ar = nullptr;
    size_t off = 0, need = 0;
    DevBuf() = default;
    // arena-backed; the caller holds the device mutex and synchronises its streams before this object dies
    explicit DevBuf(Arena* a) { if (a && !a->busy) { ar = a; ar->busy = true; } }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() {
        for (void* p : ptrs) cudaFree(p);
        if (ar) {
            cudaDeviceSynchronize();                      // what cudaFree used to imply: nothing in flight reads the arena
            if (need > ar->cap) {                         // next call of this size fits
                cudaFree(ar->base);
                ar->base = nullptr; ar->cap = 0;
                void* p = nullptr;
                const size_t want = need + need / 8;
                if (cudaMalloc(&p, want) == cudaSuccess) { ar->base = (char*)p; ar->cap = want; }
                else cudaGetLastError();
            }
            ar->busy = false;
        }
    }
    template <class T> cudaError_t alloc(T** out, size_t count) {
        const size_t bytes = (std::max<size_t>(1, count) * sizeof(T) + 255) & ~(size_t)255;
        need += bytes;
        if (ar && off + bytes <= ar->cap) { *out = (T*)(ar->base + off); off += bytes; return cudaSuccess; }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) { ptrs.push_back(p); *out = (T*)p; }
        return e;
    }
};

constexpr size_t RAW_LEN = 739624, BINCODE_LEN = 1001905;
constexpr size_t RAW_ROOTS = 8, RAW_G1 = 131080, RAW_G2 = 720904;
constexpr size_t BC_EXPANDED = 40, BC_REVERSE = 131152, BC_ROOTS = 262264, BC_G1 = 393344, BC_G2 = 983176;
constexpr int N_G2 = 65;

TableGeom make_geom(int c) {
    TableGeom g;
    g.c = c;
    g.W = (255 + c - 1) / c;
    g.half = 1u << (c - 1);
    // top digit = (r - 1) >> c(W-1), plus the carry of the signed digit below it
    const int sh = c * (g.W - 1);
    // r as 32-bit words (little endian)
    uint32_t top = 0;
    for (int bit = 0; bit < 32 && sh + bit < 256; bit++) {
        int b = sh + bit;
        uint32_t w = FR_MOD_W32::at(b >> 5);
        top |= ((w >> (b & 31)) & 1u) << bit;
    }
    g.top_entries = top + 1;
    g.per_point = (uint32_t)(g.W - 1) * g.half + g.top_entries;
    return g;
}

struct KernelTimer {   // optional CUDA-event pair around a launch
    cudaEvent_t a = nullptr, b = nullptr;
    int kind = 0;
};

struct ChunkSlot {
    uint8_t* d_blobs = nullptr;     // chunk * 131072 (host-input path only)
    uint8_t* d_q = nullptr;         // chunk * 131072 quotient scalars
    G1Xyzz* d_partials = nullptr;   // one XYZZ partial sum per MSM warp: max(chunk, 128 * 32)
    uint8_t* d_out = nullptr;       // per blob: C48 | vh32 | x32 | y32 | proof48 | hash32 | status1(+pad)
    uint8_t* d_zin = nullptr;       // chunk * 32 (compute_kzg_proof inputs)
    uint8_t* d_cin = nullptr;       // chunk * 48 (compute_blob_kzg_proof inputs)
    Fr* d_inv = nullptr;            // chunk * 4096 Fr: batch-inversion scratch of k_fr_eval_quot
    uint32_t* d_bad = nullptr;      // chunk
    uint8_t* h_out = nullptr;       // pinned mirror of d_out
    cudaEvent_t ev_in = nullptr, ev_sha = nullptr, ev_done = nullptr, ev_out = nullptr;
    bool busy = false;
    size_t first = 0, count = 0;    // blobs of the batch this slot currently holds
};

constexpr int OUT_STRIDE = 224;     // 48+32+32+32+48+32 = 224
constexpr int OFF_C = 0, OFF_VH = 48, OFF_X = 80, OFF_Y = 112, OFF_PROOF = 144, OFF_HASH = 192;

struct DeviceCtx {
    int dev = 0;
    TableGeom geom{};
    TableEntry* table = nullptr;
    uint64_t table_bytes = 0;
    Fr* roots = nullptr;
    G1Affine* g1_aff = nullptr;            // 4096 setup points, affine Montgomery
    LineStep* g2_lines = nullptr;          // device 0 only: Miller-loop lines of [s]G2 and the G2 generator
    int pairing_lanes = 2;                 // RAIKO_KZG_PAIRING_LANES: 2 = CTA-wide (144 threads), 1 = one warp (12 lanes), 0 = single thread
    cudaStream_t s_main = nullptr, s_sha = nullptr, s_in = nullptr, s_out = nullptr;
    int chunk = 0;
    int max_partials = 0;
    ChunkSlot slot[2];
    Arena arena;                           // cached per-call scratch (DevBuf)
    std::mutex mu;
    int sm_count = 148;
    int warps_per_sm = 8;
    // Blob hashing runs on the main stream ahead of the commitment MSM.  Beside it (own stream,
    // RAIKO_KZG_SHA_SERIAL=0) its latency-bound warps sit in the MSM's issue slots and cost the
    // MSM 3-6 %, more than the 3.3 ms per chunk the hash takes alone.
    bool sha_serial = true;
    bool sha_duo = true;                   // two-warp hash (schedule / rounds split); RAIKO_KZG_SHA_DUO=0: one warp
    // MSM formulation (RAIKO_KZG_MSM_AFFINE): 1 = batched affine additions (k_msm_affine) where the
    // planner estimates them faster, i.e. launches whose lanes own a few hundred table entries;
    // 0 = XYZZ only (k_msm); 2 = affine wherever eligible (tests).
    int msm_affine = 1;
    int aff_chains = MSM_AFF_MAX_K;
    int aff_warps = MSM_AFF_THREADS / 32;      // warps per CTA of the affine kernel (one CTA per SM)
    int aff_min_entries = 8 * MSM_AFF_MAX_K;   // per-lane table entries below which k_msm is used
    int max_splits_log2 = 7;                   // test knob: 0 forces one warp per blob
    bool window_reduced = false;               // auto-selection had to go below c = 15 for lack of free HBM
    bool wave_tail = true;                     // wave-aware chunk schedule and tail split (RAIKO_KZG_WAVE_TAIL=0: off)
    uint32_t* aff_scratch = nullptr;       // sm_count x chains x 39 words x 256 threads
    uint32_t recode_h[8] = {0};
    // RAIKO_KZG_TRACE=1: per-chunk timeline of every batch call (H2D, compute, D2H) on stderr
    bool trace = false;
    // stats
    bool stats_on = false;
    std::vector<KernelTimer> timers;
    rk_kzg_stats stats{};
};

}  // namespace

// One single-blob request waiting to be merged into a batch (see submit_single).
struct SingleReq {
    int mode = 0;
    const uint8_t* blob = nullptr;
    const uint8_t* aux32 = nullptr;        // z (MODE_PROOF_AT_Z) or versioned hash
    uint8_t c[48], vh[32], x[32], y[32], proof[48];
    uint8_t status = 0;
    rk_status rc = RK_OK;
    std::string err;
    bool done = false;
};

struct rk_kzg_ctx {
    // Concurrent single-blob callers (the reference keeps up to 16 requests in flight,
    // host/src/proof.rs:121) are merged: the first caller becomes the leader and runs every
    // queued request of the same kind as ONE batch; the others wait for their result.
    std::mutex coal_mu;
    std::condition_variable coal_cv;
    std::deque<SingleReq*> coal_q;
    bool coal_leader = false;
    uint64_t coal_batches = 0, coal_requests = 0;
    std::vector<DeviceCtx*> devs;
    TableGeom geom{};
    std::vector<uint8_t> g2_be;            // 65 * 192 bytes (x.c0|x.c1|y.c0|y.c1 big-endian canonical)
};

namespace {

enum { T_MSM = 1, T_FR = 2, T_SHA = 3, T_FIN = 4 };

void timer_begin(DeviceCtx* d, cudaStream_t s, int kind) {
    d->stats.total_launches++;
    if (kind == T_MSM) d->stats.msm_launches++;
    if (!d->stats_on) return;
    KernelTimer t;
    t.kind = kind;
    cudaEventCreate(&t.a);
    cudaEventCreate(&t.b);
    cudaEventRecord(t.a, s);
    d->timers.push_back(t);
}
void timer_end(DeviceCtx* d, cudaStream_t s) {
    if (!d->stats_on) return;
    cudaEventRecord(d->timers.back().b, s);
}
void timers_collect(DeviceCtx* d) {   // call after the streams are synchronised
    for (auto& t : d->timers) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            if (t.kind == T_MSM) d->stats.msm_ms += ms;
            if (t.kind == T_FR) d->stats.fr_ms += ms;
            if (t.kind == T_SHA) d->stats.sha_ms += ms;
            if (t.kind == T_FIN) d->stats.finalize_ms += ms;
        }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    d->timers.clear();
}

void free_device(DeviceCtx* d) {
    if (!d) return;
    cudaSetDevice(d->dev);
    for (auto& s : d->slot) {
        cudaFree(s.d_blobs); cudaFree(s.d_q); cudaFree(s.d_partials); cudaFree(s.d_out);
        cudaFree(s.d_zin); cudaFree(s.d_cin); cudaFree(s.d_bad); cudaFree(s.d_inv);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_sha) cudaEventDestroy(s.ev_sha);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    cudaFree(d->table); cudaFree(d->roots); cudaFree(d->g1_aff); cudaFree(d->aff_scratch); cudaFree(d->g2_lines);
    cudaFree(d->arena.base); d->arena = Arena{};
    if (d->s_main) cudaStreamDestroy(d->s_main);
    if (d->s_sha) cudaStreamDestroy(d->s_sha);
    if (d->s_in) cudaStreamDestroy(d->s_in);
    if (d->s_out) cudaStreamDestroy(d->s_out);
    delete d;
}

// Decode the settings image into affine Montgomery G1 points on the device and the
// big-endian G2 coordinates on the host.
rk_status load_setup(DeviceCtx* d, const uint8_t* data, size_t len, std::vector<uint8_t>* g2_be) {
    CUDA_TRY(cudaMalloc(&d->g1_aff, sizeof(G1Affine) * NPTS));
    DevBuf buf;                                   // scratch is released on every return below
    int* d_err = nullptr;
    CUDA_TRY(buf.alloc(&d_err, 1));
    CUDA_TRY(cudaMemset(d_err, 0, sizeof(int)));
    uint8_t* d_in = nullptr;
    size_t g1_off = 0, g2_off = 0;
    bool ref_format = false;
    if (len == RAW_LEN) {
        if (!(data[6] == 0x10 && data[7] == 0x00 && data[0] == 0)) return fail(RK_ERR_BAD_SETTINGS, "raw settings: max_width != 4096");
        g1_off = RAW_G1; g2_off = RAW_G2; ref_format = true;
    } else if (len == BINCODE_LEN) {
        uint64_t w = 0, n1 = 0, n2 = 0;
        memcpy(&w, data, 8); memcpy(&n1, data + BC_G1, 8); memcpy(&n2, data + BC_G2, 8);
        if (w != 4096 || n1 != 4096 || n2 != N_G2 || data[len - 1] != 0)
            return fail(RK_ERR_BAD_SETTINGS, "bincode settings: unexpected header fields");
        g1_off = BC_G1 + 8; g2_off = BC_G2 + 8; ref_format = true;
    } else if (len >= 16 && memcmp(data, "RKZGTS02", 8) == 0) {
        uint32_t n1, n2;
        memcpy(&n1, data + 8, 4); memcpy(&n2, data + 12, 4);
        if (n1 != NPTS || n2 != N_G2 || len != 16 + 48ull * n1 + 192ull * n2)
            return fail(RK_ERR_BAD_SETTINGS, "compact settings: bad counts / length");
        g1_off = 16; g2_off = 16 + 48ull * n1;
    } else {
        return fail(RK_ERR_BAD_SETTINGS, "unrecognised settings image (%zu bytes)", len);
    }
    const size_t g1_bytes = ref_format ? 144ull * NPTS : 48ull * NPTS;
    CUDA_TRY(buf.alloc(&d_in, g1_bytes));
    CUDA_TRY(cudaMemcpy(d_in, data + g1_off, g1_bytes, cudaMemcpyHostToDevice));
    if (ref_format) launch_k_setup_from_ref(NPTS / 64, 64, 0, 0, d_in, NPTS, d->g1_aff, d_err);
    else launch_k_setup_decompress(NPTS / 64, 64, 0, 0, d_in, NPTS, d->g1_aff, d_err);
    CUDA_TRY(cudaGetLastError());
    int err = 0;
    CUDA_TRY(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(RK_ERR_BAD_SETTINGS, "setup G1 point %d is not a valid affine curve point", err - 1);
    if (g2_be) {
        g2_be->resize(192 * N_G2);
        if (!ref_format) {
            memcpy(g2_be->data(), data + g2_off, 192 * N_G2);
        } else {
            // 65 x (X.c0 X.c1 Y.c0 Y.c1 Z.c0 Z.c1) Montgomery limbs -> big-endian canonical x, y
            uint8_t *d_g2 = nullptr, *d_o = nullptr;
            CUDA_TRY(buf.alloc(&d_g2, 288 * N_G2));
            CUDA_TRY(buf.alloc(&d_o, 288 * N_G2));
            CUDA_TRY(cudaMemcpy(d_g2, data + g2_off, 288 * N_G2, cudaMemcpyHostToDevice));
            launch_k_fp_ref_to_be((6 * N_G2 + 63) / 64, 64, 0, 0, d_g2, 6 * N_G2, d_o);
            std::vector<uint8_t> tmp(288 * N_G2);
            CUDA_TRY(cudaMemcpy(tmp.data(), d_o, tmp.size(), cudaMemcpyDeviceToHost));
            for (int i = 0; i < N_G2; i++) {
                const uint8_t* p = tmp.data() + 288 * i;
                // Z must be (1, 0)
                bool z_ok = p[4 * 48 + 47] == 1;
                for (int k = 0; k < 47; k++) z_ok = z_ok && p[4 * 48 + k] == 0;
                for (int k = 0; k < 48; k++) z_ok = z_ok && p[5 * 48 + k] == 0;
                if (!z_ok) return fail(RK_ERR_BAD_SETTINGS, "setup G2 point %d is not affine", i);
                memcpy(g2_be->data() + 192 * i, p, 192);
            }
        }
    }
    return RK_OK;
}

rk_status build_table(DeviceCtx* d) {
    const TableGeom g = d->geom;
    const size_t entries = (size_t)NPTS * g.per_point;
    d->table_bytes = entries * sizeof(TableEntry);
    CUDA_TRY(cudaMalloc(&d->table, d->table_bytes));
    const int chains = NPTS * g.W;
    G1Xyzz *bases = nullptr, *state = nullptr, *tmp = nullptr;
    G1Affine* bases_aff = nullptr;
    DevBuf buf;
    CUDA_TRY(buf.alloc(&bases, (size_t)chains));
    CUDA_TRY(buf.alloc(&bases_aff, (size_t)chains));
    CUDA_TRY(buf.alloc(&state, (size_t)chains));
    launch_k_table_bases(NPTS / 64, 64, 0, 0, d->g1_aff, g, bases);
    launch_k_table_bases_affine((chains / 16 + 63) / 64, 64, 0, 0, bases, chains, bases_aff);
    CUDA_TRY(cudaGetLastError());
    const uint32_t emax = std::max(g.half, g.top_entries);
    int D = 256;
    while ((uint32_t)D > emax && D > TABLE_NORM_G) D >>= 1;
    CUDA_TRY(buf.alloc(&tmp, (size_t)chains * D));
    for (uint32_t d0 = 1; d0 <= emax; d0 += D) {
        launch_k_table_chain((chains + 127) / 128, 128, 0, 0, bases_aff, g, d0, D, state, tmp);
        const long long groups = (long long)chains * (D / TABLE_NORM_G);
        launch_k_table_normalize((unsigned)((groups + 127) / 128), 128, 0, 0, g, d0, D, tmp, d->table);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return RK_OK;
}

rk_status alloc_slots(DeviceCtx* d) {
    // chunk: blobs per pipeline stage = a whole number of waves of one-warp-per-blob MSM work:
    // 4736 on B200 = 2 waves of k_msm_affine (148 SMs x 16 warps) = 4 waves of k_msm (x 8 warps).
    // A chunk that is not a whole number of waves idles SMs in its last wave (1024 blobs/chunk
    // measured 13 % slower), and the per-chunk fixed costs (hash latency, finalize, launch gaps)
    // amortise over more blobs.
    // 2 slots x (blobs + quotients: 128 KiB each, inversion scratch: 144 KiB) per blob = 3.8 GB.
    d->chunk = std::max(4 * d->sm_count * d->warps_per_sm, 2 * d->sm_count * d->aff_warps);
    if (const char* e = getenv("RAIKO_KZG_L2_FETCH")) {
        // experiment knob: L2 fetch granularity for DRAM misses (32 / 64 / 128 B); table entries are random
        // 96-byte gathers, so a coarser granule than a sector over-fetches
        CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
    }
    if (const char* e = getenv("RAIKO_KZG_CHUNK")) {
        int v = atoi(e);
        if (v >= 1 && v <= 16384) d->chunk = v;
    }
    d->max_partials = std::max(16 * d->chunk, 128 * 128);
    if (const char* e = getenv("RAIKO_KZG_MSM_AFFINE")) d->msm_affine = atoi(e);
    if (const char* e = getenv("RAIKO_KZG_AFFINE_CHAINS")) d->aff_chains = std::min(MSM_AFF_MAX_K, std::max(2, atoi(e) & ~1));
    d->aff_min_entries = 8 * d->aff_chains;
    if (const char* e = getenv("RAIKO_KZG_AFFINE_MIN_ENTRIES")) d->aff_min_entries = std::max(1, atoi(e));
    if (const char* e = getenv("RAIKO_KZG_MAX_SPLITS_LOG2")) d->max_splits_log2 = std::min(7, std::max(0, atoi(e)));
    if (const char* e = getenv("RAIKO_KZG_WAVE_TAIL")) d->wave_tail = atoi(e) != 0;
    if (const char* e = getenv("RAIKO_KZG_TRACE")) d->trace = atoi(e) != 0;
    if (d->msm_affine) {
        CUDA_TRY(configure_k_msm_affine());
        CUDA_TRY(cudaMalloc(&d->aff_scratch, (size_t)d->sm_count * d->aff_chains * AFF_WORDS * (32 * d->aff_warps) * sizeof(uint32_t)));
        // H = sum_{j < W-1} 2^(c-1) * 2^(cj): adding it turns unsigned digits into signed ones
        memset(d->recode_h, 0, sizeof d->recode_h);
        for (int j = 0; j < d->geom.W - 1; j++) {
            const int bit = d->geom.c * j + d->geom.c - 1;
            if (bit < 256) d->recode_h[bit >> 5] |= 1u << (bit & 31);
        }
    }
    for (auto& s : d->slot) {
        CUDA_TRY(cudaMalloc(&s.d_blobs, (size_t)d->chunk * BLOB_BYTES));   // staging of host-pointer input
        CUDA_TRY(cudaMalloc(&s.d_q, (size_t)d->chunk * BLOB_BYTES));
        CUDA_TRY(cudaMalloc(&s.d_partials, sizeof(G1Xyzz) * (size_t)d->max_partials));
        CUDA_TRY(cudaMalloc(&s.d_out, (size_t)d->chunk * OUT_STRIDE + d->chunk));
        CUDA_TRY(cudaMalloc(&s.d_zin, (size_t)d->chunk * 32));
        CUDA_TRY(cudaMalloc(&s.d_cin, (size_t)d->chunk * 48));
        CUDA_TRY(cudaMalloc(&s.d_inv, sizeof(Fr) * (size_t)d->chunk * NPTS));
        CUDA_TRY(cudaMalloc(&s.d_bad, sizeof(uint32_t) * d->chunk));
        CUDA_TRY(cudaMallocHost(&s.h_out, (size_t)d->chunk * OUT_STRIDE + d->chunk));
        CUDA_TRY(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s.ev_sha, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
    }
    return RK_OK;
}

rk_status init_device(DeviceCtx* d, const uint8_t* settings, size_t len, int window_bits,
                      std::vector<uint8_t>* g2_be) {
    CUDA_TRY(cudaSetDevice(d->dev));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, d->dev));
    d->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("RAIKO_KZG_SHA_SERIAL")) d->sha_serial = atoi(e) != 0;
    if (const char* e = getenv("RAIKO_KZG_SHA_DUO")) d->sha_duo = atoi(e) != 0;
    d->warps_per_sm = 8;                   // k_msm: 256-thread CTAs at 248 registers
    if (prop.major < 10)
        return fail(RK_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", d->dev, prop.major, prop.minor);
    int c = window_bits;
    if (c == 0) {
        if (const char* e = getenv("RAIKO_KZG_WINDOW_BITS")) c = atoi(e);
    }
    if (c == 0) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        for (c = 15; c > 8; c--) {
            TableGeom g = make_geom(c);
            double need = (double)NPTS * g.per_point * sizeof(TableEntry) + 8e9;   // table + build scratch + chunk buffers
            if (need < 0.80 * (double)free_b) break;
        }
        if (c < 15) {
            // Not silent: a narrower window means more additions per MSM (W = ceil(255 / c)) and
            // proportionally lower throughput (DESIGN.md section 8 lists c = 13 / 14).
            d->window_reduced = true;
            fprintf(stderr, "raiko_kzg: GPU %d has %.1f GB free of %.1f GB: window table narrowed to c = %d (%d additions per scalar "
                            "instead of 17); free HBM or pass window_bits explicitly\n",
                    d->dev, free_b / 1e9, total_b / 1e9, c, (255 + c - 1) / c);
        }
    }
    if (c < 4 || c > 15) return fail(RK_ERR_ARG, "window_bits %d outside 4..15", c);
    d->geom = make_geom(c);
    // Blocking streams on purpose: they order after work already queued on the legacy default
    // stream (where PyTorch produces device-resident blobs), so a device-pointer batch never
    // reads a tensor that is still being written.  They do not serialise with each other.
    CUDA_TRY(cudaStreamCreateWithFlags(&d->s_main, cudaStreamDefault));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->s_sha, cudaStreamDefault));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->s_in, cudaStreamDefault));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->s_out, cudaStreamDefault));
    rk_status st = load_setup(d, settings, len, g2_be);
    if (st != RK_OK) return st;
    CUDA_TRY(cudaMalloc(&d->roots, sizeof(Fr) * 2 * NPTS));    // w R, then w R^2 (k_fr_eval_quot)
    launch_k_roots_brp(NPTS / 128, 128, 0, 0, d->roots);
    CUDA_TRY(cudaGetLastError());
    st = build_table(d);
    if (st != RK_OK) return st;
    return alloc_slots(d);
}

// ----------------------------------------------------------------------------------------
// Batch pipeline on one device
// ----------------------------------------------------------------------------------------
enum BatchMode { MODE_COMMIT = 0, MODE_COMMIT_PROVE = 1, MODE_PROOF_AT_Z = 2, MODE_EVAL_ONLY = 3, MODE_POINT_ONLY = 4,
                 MODE_PROVE_VH = 5 /* calc_kzg_proof: challenge from (blob, vh), then the proof */,
                 MODE_BLOB_PROOF = 6 /* compute_blob_kzg_proof: EIP-4844 challenge from (blob, commitment), then the proof */ };

struct BatchArgs {
    BatchMode mode;
    const uint8_t* blobs;        // shard base (host or device)
    bool blobs_on_device;
    const uint8_t* zs;           // MODE_PROOF_AT_Z: host, n*32
    const uint8_t* vhs;          // MODE_EVAL_ONLY / MODE_POINT_ONLY / MODE_PROVE_VH: host, n*32
    const uint8_t* cins;         // MODE_BLOB_PROOF: commitments, host or device, n*48
    size_t n;
    uint8_t *out_c, *out_vh, *out_x, *out_y, *out_proof, *status;   // host or device, may be null
    bool outs_on_device;
};

struct MsmPlan {
    int splits_log2;
    bool affine;
    double cost;      // estimated duration, in units of one XYZZ addition per lane at 8 warps/SM
};

MsmPlan plan_msm(const DeviceCtx* d, size_t nblobs) {
    // One warp per (blob, split); every warp of a launch does the same amount of work, so a
    // launch runs in whole "waves" of resident warps.  Pick the split AND the formulation that
    // minimise waves x per-warp time.  Per-warp time in units of one XYZZ addition at 8 warps/SM
    // (measured, profiles/r01/affine_ab_4_lockstep_groups.txt): k_msm costs its lane's additions plus ~8 for the
    // shuffle tree; k_msm_affine runs 16 warps/SM, each addition 1.36 units, plus ~150 for the
    // chain sums and the tree -- so it wins once a lane owns a few hundred table entries.
    const double adds_per_lane = (double)NPTS * d->geom.W / 32.0;
    MsmPlan best{0, false, 1e300};
    double best_t = 1e300;
    for (int lg = 0; lg <= d->max_splits_log2; lg++) {
        if ((nblobs << lg) > (size_t)d->max_partials) break;          // one XYZZ partial per warp
        const double warps = (double)(nblobs << lg);
        const double per_lane = adds_per_lane / (double)(1 << lg);
        const double t_x = std::ceil(warps / ((double)d->sm_count * d->warps_per_sm)) * (per_lane + 8.0);
        const int entries = ((NPTS >> lg) >> 5) * d->geom.W;
        const bool eligible = d->msm_affine && d->aff_scratch && entries >= d->aff_min_entries;
        if (!(eligible && d->msm_affine == 2) && t_x < best_t * 0.999) { best_t = t_x; best = {lg, false, t_x}; }
        if (eligible) {
            const double t_a = std::ceil(warps / ((double)d->sm_count * d->aff_warps)) * (1.36 * per_lane + 150.0);
            if (t_a < best_t * 0.999) { best_t = t_a; best = {lg, true, t_a}; }
        }
    }
    return best;
}

// One MSM launch over blobs [first, first + count) of a chunk: 1 << splits_log2 warps per blob,
// its XYZZ partial sums at partials[part_off ...).
struct MsmSeg {
    int first, count, splits_log2;
    size_t part_off;
    bool affine;
};

// Wave-aware launch plan of one chunk.  Every warp of an MSM launch does the same work, so a
// launch runs in whole waves of resident warps and a chunk that is not a whole number of waves
// idles SMs in its last one (8192 blobs = 3.46 waves of 2368 cost 4 wave times).  A chunk that
// the planner would run one warp per blob is therefore cut into its whole waves plus a
// remainder, and the remainder is planned on its own: it gets the split that fills the machine
// (1088 blobs -> 2 warps per blob = 0.92 wave of half-size work), so 3.46 waves cost 3.5.
int plan_chunk(const DeviceCtx* d, int n, MsmSeg (&seg)[2]) {
    const MsmPlan whole = plan_msm(d, (size_t)n);
    seg[0] = {0, n, whole.splits_log2, 0, whole.affine};
    // candidate: whole waves of the one-warp-per-blob affine kernel + a separately planned remainder
    const int wave = d->sm_count * d->aff_warps;
    const int entries = (NPTS >> 5) * d->geom.W;
    const bool affine_ok = d->msm_affine && d->aff_scratch && entries >= d->aff_min_entries;
    if (!d->wave_tail || !affine_ok || n <= wave || n % wave == 0) return 1;
    const int body = n / wave * wave;
    const MsmPlan rest = plan_msm(d, (size_t)(n - body));
    const double t_body = (double)(body / wave) * (1.36 * (double)NPTS * d->geom.W / 32.0 + 150.0);
    if (t_body + rest.cost >= 0.98 * whole.cost) return 1;          // nothing to gain: keep one launch
    seg[0] = {0, body, 0, 0, true};
    seg[1] = {body, n - body, rest.splits_log2, (size_t)body, rest.affine};
    return 2;
}

void launch_msm_seg(DeviceCtx* d, const uint8_t* scalars, ChunkSlot& s, uint32_t* bad, const MsmSeg& g) {
    const int n = g.count;
    const uint8_t* sc = scalars + (size_t)g.first * BLOB_BYTES;
    uint32_t* bd = bad ? bad + g.first : nullptr;
    const long long warps = (long long)n << g.splits_log2;
    timer_begin(d, d->s_main, T_MSM);
    if (g.affine) {
        MsmAffParams q;
        q.table = d->table; q.g = d->geom; q.scalars = sc; q.nblobs = n; q.splits_log2 = g.splits_log2;
        q.partials = s.d_partials + g.part_off; q.bad = bd; q.scratch = d->aff_scratch; q.K = d->aff_chains;
        const int wpc = d->aff_warps, threads = 32 * wpc;
        q.ngroups = (int)((warps + wpc - 1) / wpc);
        memcpy(q.H, d->recode_h, sizeof q.H);
        const unsigned grid = (unsigned)std::min<long long>(q.ngroups, d->sm_count);
        launch_k_msm_affine(grid, threads, (size_t)q.K * threads * 4, d->s_main, q);
        d->stats.msm_affine_launches++;
        d->stats.msm_affine_point_adds += (uint64_t)n * NPTS * d->geom.W;
    } else {
        MsmParams p;
        p.table = d->table; p.g = d->geom; p.scalars = sc; p.nblobs = n;
        p.splits_log2 = g.splits_log2;
        p.partials = s.d_partials + g.part_off; p.bad = bd;
        launch_k_msm((unsigned)((warps + 7) / 8), 256, 0, d->s_main, p);
    }
    timer_end(d, d->s_main);
    d->stats.msm_point_adds += (uint64_t)n * NPTS * d->geom.W;
}

// MSM over the chunk's scalars followed by the reduction of the partial sums to compressed
// points at out_g1 (+ versioned hashes at out_vh when given), segment by segment.
void msm_and_finalize(DeviceCtx* d, const uint8_t* scalars, int n, ChunkSlot& s, uint32_t* bad_msm, uint8_t* out_g1,
                      uint8_t* out_vh, uint8_t* o_stat) {
    MsmSeg seg[2];
    const int nseg = plan_chunk(d, n, seg);
    for (int i = 0; i < nseg; i++) launch_msm_seg(d, scalars, s, bad_msm, seg[i]);
    timer_begin(d, d->s_main, T_FIN);
    for (int i = 0; i < nseg; i++) {
        const MsmSeg& g = seg[i];
        const int splits = 1 << g.splits_log2;
        const G1Xyzz* part = s.d_partials + g.part_off;
        uint8_t* og = out_g1 + (size_t)g.first * OUT_STRIDE;
        uint8_t* ov = out_vh ? out_vh + (size_t)g.first * OUT_STRIDE : nullptr;
        if (splits >= 8) launch_k_finalize_warp(g.count, 32, 0, d->s_main, part, splits, g.count, s.d_bad + g.first, og, ov, o_stat + g.first, OUT_STRIDE);
        else launch_k_finalize((g.count + 31) / 32, 32, 0, d->s_main, part, splits, g.count, s.d_bad + g.first, og, ov, o_stat + g.first, OUT_STRIDE);
        if (i) d->stats.total_launches++;
    }
    timer_end(d, d->s_main);
}

// Chunk schedule of one shard: whole multiples of the one-warp-per-blob wave wherever possible,
// a one-wave first chunk for host input (compute starts after 310 MB of H2D instead of 620 MB),
// and whatever is left as the last chunk (plan_chunk splits its partial wave).
std::vector<size_t> chunk_schedule(const DeviceCtx* d, size_t n, bool host_input) {
    std::vector<size_t> sizes;
    const size_t chunk = (size_t)d->chunk;
    const size_t wave = (size_t)d->sm_count * (d->msm_affine ? d->aff_warps : d->warps_per_sm);
    size_t left = n;
    if (host_input && d->wave_tail && wave < chunk && n > wave) { sizes.push_back(wave); left -= wave; }
    while (left > 0) { const size_t c = std::min(chunk, left); sizes.push_back(c); left -= c; }
    return sizes;
}

rk_status run_shard(DeviceCtx* d, const BatchArgs& a) {
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    const size_t chunk = (size_t)d->chunk;
    // Whatever way this function is left, no slot stays marked busy and no copy into the caller's
    // arrays is still in flight: a failed call must not leak state into the next one.
    struct SlotReset {
        DeviceCtx* d;
        bool ok = false;
        ~SlotReset() {
            if (ok) return;
            cudaStreamSynchronize(d->s_in); cudaStreamSynchronize(d->s_main);
            cudaStreamSynchronize(d->s_sha); cudaStreamSynchronize(d->s_out);
            cudaGetLastError();
            for (auto& s : d->slot) { s.busy = false; s.count = 0; }
            timers_collect(d);
        }
    } reset{d};
    auto drain = [&](ChunkSlot& s) -> rk_status {
        if (!s.busy) return RK_OK;
        CUDA_TRY(cudaEventSynchronize(s.ev_out));
        s.busy = false;
        if (a.outs_on_device) return RK_OK;
        const uint8_t* stat = s.h_out + (size_t)d->chunk * OUT_STRIDE;
        for (size_t i = 0; i < s.count; i++) {
            const uint8_t* o = s.h_out + i * OUT_STRIDE;
            const size_t gi = s.first + i;
            if (a.out_c) memcpy(a.out_c + 48 * gi, o + OFF_C, 48);
            if (a.out_vh) memcpy(a.out_vh + 32 * gi, o + OFF_VH, 32);
            if (a.out_x) memcpy(a.out_x + 32 * gi, o + OFF_X, 32);
            if (a.out_y) memcpy(a.out_y + 32 * gi, o + OFF_Y, 32);
            if (a.out_proof) memcpy(a.out_proof + 48 * gi, o + OFF_PROOF, 48);
            if (a.status) a.status[gi] = stat[i];
        }
        return RK_OK;
    };

    const std::vector<size_t> sizes = chunk_schedule(d, a.n, !a.blobs_on_device);
    // optional timeline: five timed events per chunk (H2D begin / end, compute begin / end, D2H end)
    struct ChunkTrace { cudaEvent_t e[5]; int cnt; int nseg; };
    std::vector<ChunkTrace> tr;
    cudaEvent_t tr0 = nullptr;
    if (d->trace) {
        tr.resize(sizes.size());
        for (auto& t : tr) for (auto& e : t.e) cudaEventCreate(&e);
        cudaEventCreate(&tr0);
        cudaEventRecord(tr0, d->s_in);
    }
    size_t first = 0;
    for (size_t ci = 0; ci < sizes.size(); ci++) {
        ChunkSlot& s = d->slot[ci & 1];
        rk_status st = drain(s);
        if (st != RK_OK) return st;
        const int cnt = (int)sizes[ci];
        if (d->trace) { tr[ci].cnt = cnt; cudaEventRecord(tr[ci].e[0], d->s_in); }

        // ---- input ---------------------------------------------------------------------
        const uint8_t* d_blobs;
        if (a.blobs_on_device) {
            d_blobs = a.blobs + first * BLOB_BYTES;
        } else {
            CUDA_TRY(cudaMemcpyAsync(s.d_blobs, a.blobs + first * BLOB_BYTES, (size_t)cnt * BLOB_BYTES,
                                     cudaMemcpyHostToDevice, d->s_in));
            d->stats.h2d_bytes += (uint64_t)cnt * BLOB_BYTES;
            d_blobs = s.d_blobs;
        }
        if (a.mode == MODE_PROOF_AT_Z) {
            CUDA_TRY(cudaMemcpyAsync(s.d_zin, a.zs + 32 * first, 32 * (size_t)cnt, cudaMemcpyHostToDevice, d->s_in));
            d->stats.h2d_bytes += 32ull * cnt;
        }
        if (a.mode == MODE_BLOB_PROOF) {
            CUDA_TRY(cudaMemcpyAsync(s.d_cin, a.cins + 48 * first, 48 * (size_t)cnt, cudaMemcpyDefault, d->s_in));
            d->stats.h2d_bytes += 48ull * cnt;
        }
        uint8_t* o = s.d_out;
        uint8_t* o_stat = s.d_out + chunk * OUT_STRIDE;
        if (a.mode == MODE_EVAL_ONLY || a.mode == MODE_POINT_ONLY || a.mode == MODE_PROVE_VH) {
            // caller-supplied versioned hashes go where the commit stage would have put them
            CUDA_TRY(cudaMemcpy2DAsync(o + OFF_VH, OUT_STRIDE, a.vhs + 32 * first, 32, 32, (size_t)cnt,
                                       cudaMemcpyHostToDevice, d->s_in));
        }
        if (d->trace) cudaEventRecord(tr[ci].e[1], d->s_in);
        CUDA_TRY(cudaEventRecord(s.ev_in, d->s_in));
        CUDA_TRY(cudaStreamWaitEvent(d->s_main, s.ev_in, 0));
        if (d->trace) cudaEventRecord(tr[ci].e[2], d->s_main);
        CUDA_TRY(cudaMemsetAsync(s.d_bad, 0, sizeof(uint32_t) * cnt, d->s_main));
        CUDA_TRY(cudaMemsetAsync(o_stat, 0, (size_t)cnt, d->s_main));

        const bool need_hash = a.mode == MODE_COMMIT_PROVE || a.mode == MODE_EVAL_ONLY || a.mode == MODE_POINT_ONLY || a.mode == MODE_PROVE_VH;
        if (need_hash) {
            // large chunks: hash first on the main stream; small (latency-bound) batches: beside the MSM
            // (never beside the MSM when the call pipelines several chunks: the hash of chunk i+1 then
            // lands between the kernels of chunk i and holds SMs the persistent MSM kernel needs --
            // measured 23 ms lost per 4736-blob chunk, profiles/r02/e2e_trace_8192.txt)
            const bool serial = d->sha_serial && (cnt >= d->sm_count * d->warps_per_sm || sizes.size() > 1);
            cudaStream_t hs = serial ? d->s_main : d->s_sha;
            if (!serial) CUDA_TRY(cudaStreamWaitEvent(d->s_sha, s.ev_in, 0));
            timer_begin(d, hs, T_SHA);
            // Two-warp hash (lane = blob; schedule and rounds on separate warps) for every batch size.
            if (d->sha_duo) launch_k_sha_blob_duo((cnt + 31) / 32, 64, 0, hs, d_blobs, cnt, o + OFF_HASH, OUT_STRIDE);
            else launch_k_sha_blob((cnt + 31) / 32, 32, 0, hs, d_blobs, cnt, o + OFF_HASH, OUT_STRIDE);
            timer_end(d, hs);
            CUDA_TRY(cudaEventRecord(s.ev_sha, hs));
        }
        // ---- commitment ------------------------------------------------------------------
        if (a.mode == MODE_COMMIT || a.mode == MODE_COMMIT_PROVE)
            msm_and_finalize(d, d_blobs, cnt, s, s.d_bad, o + OFF_C, o + OFF_VH, o_stat);
        // ---- evaluation / quotient -------------------------------------------------------
        if (a.mode != MODE_COMMIT) {
            if (need_hash) CUDA_TRY(cudaStreamWaitEvent(d->s_main, s.ev_sha, 0));
            if (a.mode == MODE_BLOB_PROOF) {       // z = compute_challenge(blob, commitment), Deneb spec
                timer_begin(d, d->s_main, T_SHA);
                launch_k_sha_fs_challenge((cnt + 31) / 32, 32, 0, d->s_main, d_blobs, s.d_cin, cnt, s.d_zin);
                timer_end(d, d->s_main);
            }
            FrParams fp;
            fp.blobs = d_blobs; fp.roots_brp = d->roots;
            fp.blob_hash = o + OFF_HASH; fp.vh = o + OFF_VH; fp.z_in = s.d_zin;
            fp.mode = (a.mode == MODE_PROOF_AT_Z || a.mode == MODE_BLOB_PROOF) ? 1 : 0;
            fp.want_quotient = (a.mode == MODE_COMMIT_PROVE || a.mode == MODE_PROOF_AT_Z || a.mode == MODE_PROVE_VH || a.mode == MODE_BLOB_PROOF) ? 1 : 0;
            fp.eval = (a.mode == MODE_POINT_ONLY) ? 0 : 1;
            fp.nblobs = cnt; fp.out_x = o + OFF_X; fp.out_y = o + OFF_Y; fp.q_out = s.d_q; fp.bad = s.d_bad;
            fp.out_stride = OUT_STRIDE;
            fp.inv_scratch = s.d_inv;
            timer_begin(d, d->s_main, T_FR);
            launch_k_fr_eval_quot(cnt, FR_THREADS, FR_SMEM_BYTES, d->s_main, fp);
            timer_end(d, d->s_main);
            if (fp.want_quotient) msm_and_finalize(d, s.d_q, cnt, s, nullptr, o + OFF_PROOF, nullptr, o_stat);
            launch_k_status_only((cnt + 127) / 128, 128, 0, d->s_main, s.d_bad, cnt, o, o_stat, OUT_STRIDE, OFF_HASH);
            d->stats.total_launches++;
        }
        CUDA_TRY(cudaGetLastError());
        if (d->trace) cudaEventRecord(tr[ci].e[3], d->s_main);
        CUDA_TRY(cudaEventRecord(s.ev_done, d->s_main));
        // ---- output ----------------------------------------------------------------------
        CUDA_TRY(cudaStreamWaitEvent(d->s_out, s.ev_done, 0));
        if (a.outs_on_device) {
            auto scatter = [&](uint8_t* dst, int off, int width) -> cudaError_t {
                if (!dst) return cudaSuccess;
                return cudaMemcpy2DAsync(dst + (size_t)width * first, width, o + off, OUT_STRIDE, width, (size_t)cnt,
                                         cudaMemcpyDeviceToDevice, d->s_out);
            };
            CUDA_TRY(scatter(a.out_c, OFF_C, 48));
            CUDA_TRY(scatter(a.out_vh, OFF_VH, 32));
            CUDA_TRY(scatter(a.out_x, OFF_X, 32));
            CUDA_TRY(scatter(a.out_y, OFF_Y, 32));
            CUDA_TRY(scatter(a.out_proof, OFF_PROOF, 48));
            if (a.status) CUDA_TRY(cudaMemcpyAsync(a.status + first, o_stat, (size_t)cnt, cudaMemcpyDeviceToDevice, d->s_out));
        } else {
            CUDA_TRY(cudaMemcpyAsync(s.h_out, o, (size_t)cnt * OUT_STRIDE, cudaMemcpyDeviceToHost, d->s_out));
            CUDA_TRY(cudaMemcpyAsync(s.h_out + chunk * OUT_STRIDE, o_stat, (size_t)cnt, cudaMemcpyDeviceToHost, d->s_out));
            d->stats.d2h_bytes += (uint64_t)cnt * (OUT_STRIDE + 1);
        }
        if (d->trace) cudaEventRecord(tr[ci].e[4], d->s_out);
        CUDA_TRY(cudaEventRecord(s.ev_out, d->s_out));
        // only now does the slot hold results that drain() may copy out
        s.first = first; s.count = (size_t)cnt; s.busy = true;
        first += (size_t)cnt;
    }
    for (auto& s : d->slot) {
        rk_status st = drain(s);
        if (st != RK_OK) return st;
    }
    CUDA_TRY(cudaStreamSynchronize(d->s_main));
    CUDA_TRY(cudaStreamSynchronize(d->s_sha));
    CUDA_TRY(cudaStreamSynchronize(d->s_out));
    timers_collect(d);
    if (d->trace) {
        CUDA_TRY(cudaStreamSynchronize(d->s_in));
        std::string line;
        char buf[256];
        snprintf(buf, sizeof buf, "[raiko_kzg trace] GPU %d mode %d n %zu %s input, %zu chunks; ms from the first H2D enqueue: blobs | h2d begin..end | compute begin..end | d2h end\n",
                 d->dev, (int)a.mode, a.n, a.blobs_on_device ? "device" : "host", sizes.size());
        line += buf;
        for (size_t ci = 0; ci < tr.size(); ci++) {
            float t[5];
            for (int k = 0; k < 5; k++) { t[k] = 0.f; cudaEventElapsedTime(&t[k], tr0, tr[ci].e[k]); }
            snprintf(buf, sizeof buf, "[raiko_kzg trace]   chunk %2zu  %5d | %8.2f .. %8.2f | %8.2f .. %8.2f | %8.2f\n", ci, tr[ci].cnt, t[0], t[1], t[2], t[3], t[4]);
            line += buf;
        }
        fputs(line.c_str(), stderr);
        for (auto& t : tr) for (auto& e : t.e) cudaEventDestroy(e);
        cudaEventDestroy(tr0);
    }
    reset.ok = true;
    return RK_OK;
}

// Device that owns `p`, or -1 for host memory (pageable or pinned) and NULL.
int device_of_ptr(const void* p) {
    if (!p) return -1;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return attr.device;
    return -1;
}
bool is_device_ptr(const void* p) { return device_of_ptr(p) >= 0; }
// A device pointer handed to this library must live on the context's first device, as the header
// promises; anything else would be read through peer access at best and fault at worst.
rk_status check_ptr_device(const rk_kzg_ctx* ctx, const void* p, const char* what) {
    const int dev = device_of_ptr(p);
    if (dev >= 0 && dev != ctx->devs[0]->dev)
        return fail(RK_ERR_ARG, "%s is device memory of GPU %d but the context's first device is GPU %d", what, dev, ctx->devs[0]->dev);
    return RK_OK;
}

rk_status run_batch(rk_kzg_ctx* ctx, BatchArgs a) {
    if (!ctx) return fail(RK_ERR_ARG, "null context");
    if (a.n == 0) return RK_OK;
    if (!a.blobs) return fail(RK_ERR_ARG, "null blobs pointer");
    DeviceGuard dg;
    cudaSetDevice(ctx->devs[0]->dev);
    a.blobs_on_device = is_device_ptr(a.blobs);
    rk_status pst = check_ptr_device(ctx, a.blobs, "blobs");
    if (pst != RK_OK) return pst;
    uint8_t* outs[6] = {a.out_c, a.out_vh, a.out_x, a.out_y, a.out_proof, a.status};
    int n_dev_out = 0, n_out = 0;
    for (uint8_t* p : outs) if (p) {
        n_out++; n_dev_out += is_device_ptr(p) ? 1 : 0;
        if ((pst = check_ptr_device(ctx, p, "an output array")) != RK_OK) return pst;
    }
    if (n_dev_out != 0 && n_dev_out != n_out) return fail(RK_ERR_ARG, "outputs must be all host or all device pointers");
    a.outs_on_device = n_dev_out != 0;
    if ((a.blobs_on_device || a.outs_on_device) && ctx->devs.size() != 1)
        return fail(RK_ERR_ARG, "device pointers need a single-device context");
    if (a.blobs_on_device && ((uintptr_t)a.blobs & 15)) return fail(RK_ERR_ARG, "device blob pointer must be 16-byte aligned");
    const size_t ndev = ctx->devs.size();
    if (ndev == 1 || a.n < ndev) return run_shard(ctx->devs[0], a);
    // contiguous shards, one host thread per device, host-side gather (no collective)
    std::vector<rk_status> rc(ndev, RK_OK);
    std::vector<std::string> msg(ndev);
    std::vector<std::thread> th;
    for (size_t g = 0; g < ndev; g++) {
        const size_t lo = a.n * g / ndev, hi = a.n * (g + 1) / ndev;
        BatchArgs s = a;
        s.blobs = a.blobs + lo * BLOB_BYTES;
        s.n = hi - lo;
        if (a.zs) s.zs = a.zs + 32 * lo;
        if (a.vhs) s.vhs = a.vhs + 32 * lo;
        if (a.cins) s.cins = a.cins + 48 * lo;
        if (a.out_c) s.out_c = a.out_c + 48 * lo;
        if (a.out_vh) s.out_vh = a.out_vh + 32 * lo;
        if (a.out_x) s.out_x = a.out_x + 32 * lo;
        if (a.out_y) s.out_y = a.out_y + 32 * lo;
        if (a.out_proof) s.out_proof = a.out_proof + 48 * lo;
        if (a.status) s.status = a.status + lo;
        th.emplace_back([&, g, s]() { rc[g] = run_shard(ctx->devs[g], s); if (rc[g] != RK_OK) msg[g] = g_last_error; });
    }
    for (auto& t : th) t.join();
    for (size_t g = 0; g < ndev; g++)
        if (rc[g] != RK_OK) { g_last_error = msg[g]; return rc[g]; }
    return RK_OK;
}

// ----------------------------------------------------------------------------------------
// Verification (single device: ctx device 0)
// ----------------------------------------------------------------------------------------
constexpr size_t VERIFY_MAX_N = 16384;

// Scratch and decompression stage of one verification call.  verify_begin() allocates, then
// decompresses / subgroup-checks the commitments and proofs on the side stream (s_sha) -- they do
// not depend on z or y, so they overlap the challenge / evaluation kernels of the blob path and the
// transcript hash.  verify_finish() runs the rest on the main stream and joins.
struct VerifyState {
    int n = 0;
    const uint8_t *d_c = nullptr, *d_p = nullptr;
    G1Affine *d_pts = nullptr, *d_pts_hi = nullptr, *d_pair = nullptr;   // d_pts_hi: 2^128 * (C_i | proof_i | G)
    int *d_inf = nullptr, *d_pinf = nullptr, *d_flags = nullptr;
    G1Xyzz *d_a = nullptr, *d_e = nullptr, *d_part = nullptr;
    Fr* d_r = nullptr;
    uint8_t* d_g2 = nullptr;
    cudaEvent_t ev_inputs = nullptr, ev_points = nullptr;
    ~VerifyState() { if (ev_inputs) cudaEventDestroy(ev_inputs); if (ev_points) cudaEventDestroy(ev_points); }
};
constexpr int VR_PARTS = 64;               // CTAs of the first reduction level

// d_c, d_p: n x 48 compressed, on the device; the copies that produced them were enqueued on s_main.
rk_status verify_begin(rk_kzg_ctx* ctx, DeviceCtx* d, DevBuf& buf, VerifyState& v, int n, const uint8_t* d_c, const uint8_t* d_p) {
    v.n = n; v.d_c = d_c; v.d_p = d_p;
    CUDA_TRY(buf.alloc(&v.d_pts, 2 * (size_t)n)); CUDA_TRY(buf.alloc(&v.d_inf, 2 * (size_t)n));
    CUDA_TRY(buf.alloc(&v.d_pts_hi, 2 * (size_t)n + 1));
    CUDA_TRY(buf.alloc(&v.d_a, 2 * (size_t)n)); CUDA_TRY(buf.alloc(&v.d_e, 6 * (size_t)n)); CUDA_TRY(buf.alloc(&v.d_part, 2 * (size_t)VR_PARTS));
    CUDA_TRY(buf.alloc(&v.d_r, 1)); CUDA_TRY(buf.alloc(&v.d_pair, 2)); CUDA_TRY(buf.alloc(&v.d_pinf, 2));
    CUDA_TRY(buf.alloc(&v.d_flags, 4)); CUDA_TRY(buf.alloc(&v.d_g2, 384));
    CUDA_TRY(cudaEventCreateWithFlags(&v.ev_inputs, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&v.ev_points, cudaEventDisableTiming));
    cudaStream_t st = d->s_main, side = d->s_sha;
    CUDA_TRY(cudaMemsetAsync(v.d_flags, 0, 4 * sizeof(int), st));
    // [s]G2 then the generator (only the single-thread pairing kernel reads them)
    CUDA_TRY(cudaMemcpyAsync(v.d_g2, ctx->g2_be.data() + 192, 192, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(v.d_g2 + 192, ctx->g2_be.data(), 192, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(v.ev_inputs, st));
    CUDA_TRY(cudaStreamWaitEvent(side, v.ev_inputs, 0));
    launch_k_g1_decompress_validate((n + 31) / 32, 32, 0, side, d_c, n, v.d_pts, v.d_inf, v.d_flags + 0);
    launch_k_g1_decompress_validate((n + 31) / 32, 32, 0, side, d_p, n, v.d_pts + n, v.d_inf + n, v.d_flags + 1);
    // 2^128 * every point (and * G): the base points of the scalars' upper halves, independent of the challenge
    launch_k_verify_pow128((2 * n + 1 + 63) / 64, 64, 0, side, v.d_pts, v.d_inf, 2 * n, v.d_pts_hi);
    CUDA_TRY(cudaEventRecord(v.ev_points, side));
    d->stats.total_launches += 3;
    return RK_OK;
}

// d_z, d_y: n x 32 big-endian canonical, on the device, produced on s_main.
rk_status verify_finish(rk_kzg_ctx* ctx, DeviceCtx* d, VerifyState& v, const uint8_t* d_z, const uint8_t* d_y, int* out_ok) {
    (void)ctx;
    const int n = v.n;
    cudaStream_t st = d->s_main;
    // RAIKO_KZG_VERIFY_TRACE=1 prints the per-stage device time of this call to stderr
    const bool trace = getenv("RAIKO_KZG_VERIFY_TRACE") != nullptr;
    cudaEvent_t ev[6];
    if (trace) for (auto& e : ev) cudaEventCreate(&e);
    auto mark = [&](int i) { if (trace) cudaEventRecord(ev[i], st); };
    mark(0);
    launch_k_batch_challenge(1, 64, 0, st, v.d_c, d_z, d_y, v.d_p, n, v.d_r);
    mark(1);
    CUDA_TRY(cudaStreamWaitEvent(st, v.ev_points, 0));
    mark(2);
    launch_k_verify_terms(dim3((unsigned)((n + 63) / 64), 4, 2), 64, 0, st, v.d_r, d_z, d_y, v.d_pts, v.d_inf, v.d_pts + n, v.d_inf + n, n,
                          v.d_pts_hi, v.d_pts_hi + n, v.d_pts_hi + 2 * n, v.d_a, v.d_e, v.d_flags + 2);
    mark(3);
    launch_k_verify_reduce_partial(VR_PARTS, VR_THREADS, 0, st, v.d_a, 2 * n, v.d_e, 6 * n, v.d_part, v.d_part + VR_PARTS);
    launch_k_verify_reduce(1, VR_THREADS, 0, st, v.d_part, v.d_part + VR_PARTS, VR_PARTS, v.d_pair, v.d_pinf);
    mark(4);
    if (d->pairing_lanes >= 2 && d->g2_lines) launch_k_pairing_check_cta(1, PAIRING_CTA_THREADS, 0, st, v.d_pair, v.d_pinf, d->g2_lines, v.d_flags + 3);
    else if (d->pairing_lanes && d->g2_lines) launch_k_pairing_check_lanes(1, 32, 0, st, v.d_pair, v.d_pinf, d->g2_lines, v.d_flags + 3);
    else launch_k_pairing_check(1, 1, 0, st, v.d_pair, v.d_pinf, v.d_g2, v.d_g2 + 192, v.d_flags + 3);
    mark(5);
    d->stats.total_launches += 5;
    CUDA_TRY(cudaGetLastError());
    int flags[4];
    CUDA_TRY(cudaMemcpyAsync(flags, v.d_flags, sizeof flags, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaStreamSynchronize(d->s_sha));
    if (trace) {
        const char* names[5] = {"transcript hash", "wait for decompress+subgroup (side stream)", "r-power terms", "reduce", "pairing"};
        for (int i = 0; i < 5; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); fprintf(stderr, "[verify n=%d] %-44s %8.3f ms\n", n, names[i], ms); }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    if (flags[0]) return fail(RK_ERR_BAD_POINT, "commitment %d is not a valid G1 point", flags[0] - 1);
    if (flags[1]) return fail(RK_ERR_BAD_POINT, "proof %d is not a valid G1 point", flags[1] - 1);
    if (flags[2]) return fail(RK_ERR_NONCANONICAL_FE, "z or y of element %d is not a canonical field element", flags[2] - 1);
    *out_ok = flags[3];
    return RK_OK;
}

rk_status verify_core(rk_kzg_ctx* ctx, DeviceCtx* d, DevBuf& buf, int n, const uint8_t* d_c, const uint8_t* d_p,
                      const uint8_t* d_z, const uint8_t* d_y, int* out_ok) {
    VerifyState v;
    rk_status st = verify_begin(ctx, d, buf, v, n, d_c, d_p);
    if (st != RK_OK) { cudaStreamSynchronize(d->s_main); cudaStreamSynchronize(d->s_sha); return st; }
    st = verify_finish(ctx, d, v, d_z, d_y, out_ok);
    if (st != RK_OK) { cudaStreamSynchronize(d->s_main); cudaStreamSynchronize(d->s_sha); }
    return st;
}

rk_status verify_blob_batch_device(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs,
                                   size_t n, int* out_ok) {
    DeviceCtx* d = ctx->devs[0];
    DeviceGuard dg;
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    const bool blobs_dev = is_device_ptr(blobs);
    if (blobs_dev && ((uintptr_t)blobs & 15)) return fail(RK_ERR_ARG, "device blob pointer must be 16-byte aligned");
    for (const void* p : {(const void*)blobs, (const void*)commitments, (const void*)proofs}) {
        rk_status pst = check_ptr_device(ctx, p, "an input array");
        if (pst != RK_OK) return pst;
    }
    DevBuf buf(&d->arena);
    uint8_t *d_c = nullptr, *d_p = nullptr, *d_z = nullptr, *d_y = nullptr;
    uint32_t* d_bad = nullptr;
    CUDA_TRY(buf.alloc(&d_c, 48 * n)); CUDA_TRY(buf.alloc(&d_p, 48 * n));
    CUDA_TRY(buf.alloc(&d_z, 32 * n)); CUDA_TRY(buf.alloc(&d_y, 32 * n)); CUDA_TRY(buf.alloc(&d_bad, n));
    cudaStream_t st = d->s_main;
    CUDA_TRY(cudaMemcpyAsync(d_c, commitments, 48 * n, cudaMemcpyDefault, st));
    CUDA_TRY(cudaMemcpyAsync(d_p, proofs, 48 * n, cudaMemcpyDefault, st));
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t) * n, st));
    VerifyState v;                          // declared after buf: destroyed first
    {
        rk_status vst = verify_begin(ctx, d, buf, v, (int)n, d_c, d_p);     // decompression runs beside the loop below
        if (vst != RK_OK) { cudaStreamSynchronize(st); cudaStreamSynchronize(d->s_sha); return vst; }
    }
    struct Join { DeviceCtx* d; ~Join() { cudaStreamSynchronize(d->s_main); cudaStreamSynchronize(d->s_sha); cudaStreamSynchronize(d->s_in); } } join{d};
    const size_t chunk = (size_t)d->chunk;
    for (size_t first = 0, ci = 0; first < n; first += chunk, ci++) {
        ChunkSlot& s = d->slot[ci & 1];
        const int cnt = (int)std::min(chunk, n - first);
        const uint8_t* d_blobs;
        if (blobs_dev) {
            d_blobs = blobs + first * BLOB_BYTES;
        } else {
            if (ci >= 2) CUDA_TRY(cudaEventSynchronize(s.ev_done));        // slot free again?
            CUDA_TRY(cudaMemcpyAsync(s.d_blobs, blobs + first * BLOB_BYTES, (size_t)cnt * BLOB_BYTES, cudaMemcpyHostToDevice, d->s_in));
            CUDA_TRY(cudaEventRecord(s.ev_in, d->s_in));
            CUDA_TRY(cudaStreamWaitEvent(st, s.ev_in, 0));
            d->stats.h2d_bytes += (uint64_t)cnt * BLOB_BYTES;
            d_blobs = s.d_blobs;
        }
        timer_begin(d, st, T_SHA);
        launch_k_sha_fs_challenge((cnt + 31) / 32, 32, 0, st, d_blobs, d_c + 48 * first, cnt, d_z + 32 * first);
        timer_end(d, st);
        FrParams fp{};
        fp.blobs = d_blobs; fp.roots_brp = d->roots; fp.z_in = d_z + 32 * first; fp.mode = 1; fp.want_quotient = 0; fp.eval = 1;
        fp.nblobs = cnt; fp.out_stride = 32; fp.out_x = nullptr; fp.out_y = d_y + 32 * first; fp.q_out = nullptr; fp.bad = d_bad + first;
        fp.inv_scratch = s.d_inv;
        timer_begin(d, st, T_FR);
        launch_k_fr_eval_quot(cnt, FR_THREADS, FR_SMEM_BYTES, st, fp);
        timer_end(d, st);
        CUDA_TRY(cudaEventRecord(s.ev_done, st));
    }
    CUDA_TRY(cudaGetLastError());
    std::vector<uint32_t> bad(n);
    CUDA_TRY(cudaMemcpyAsync(bad.data(), d_bad, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    timers_collect(d);
    for (size_t i = 0; i < n; i++)
        if (bad[i]) return fail(RK_ERR_NONCANONICAL_FE, "blob %zu: Failed to deserialize blob to field elements", i);
    return verify_finish(ctx, d, v, d_z, d_y, out_ok);
}

rk_status check_blob_len(size_t len) {
    if (len != BLOB_BYTES) return fail(RK_ERR_BAD_LENGTH, "blob length %zu != %d", len, BLOB_BYTES);
    return RK_OK;
}

}  // namespace

// ========================================================================================
// C ABI
// ========================================================================================
extern "C" {

const char* rk_last_error(void) { return g_last_error.c_str(); }
const char* rk_version(void) { return "raiko_b200-kzg 0.1 (sm_100a, 30-bit-limb IMAD.WIDE, fixed-base window table)"; }

rk_status rk_kzg_ctx_create_ex(const uint8_t* settings, size_t len, const int* devices, int ndev, int window_bits,
                               rk_kzg_ctx** out) {
    if (!settings || !out) return fail(RK_ERR_ARG, "null argument");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(RK_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    std::vector<int> devs;
    if (!devices || ndev <= 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        devs.push_back(cur);
    } else {
        for (int i = 0; i < ndev; i++) {
            if (devices[i] < 0 || devices[i] >= count) return fail(RK_ERR_ARG, "device ordinal %d out of range", devices[i]);
            devs.push_back(devices[i]);
        }
    }
    int prev = 0;
    cudaGetDevice(&prev);
    auto* ctx = new rk_kzg_ctx;
    std::vector<rk_status> rc(devs.size(), RK_OK);
    std::vector<std::string> msg(devs.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < devs.size(); i++) {
        auto* d = new DeviceCtx;
        d->dev = devs[i];
        ctx->devs.push_back(d);
    }
    for (size_t i = 0; i < devs.size(); i++)
        th.emplace_back([&, i]() {
            rc[i] = init_device(ctx->devs[i], settings, len, window_bits, i == 0 ? &ctx->g2_be : nullptr);
            if (rc[i] != RK_OK) msg[i] = g_last_error;
        });
    for (auto& t : th) t.join();
    cudaSetDevice(prev);
    for (size_t i = 0; i < devs.size(); i++)
        if (rc[i] != RK_OK) {
            g_last_error = msg[i];
            rk_status st = rc[i];
            rk_kzg_ctx_destroy(ctx);
            return st;
        }
    ctx->geom = ctx->devs[0]->geom;
    {
        // fixed-argument pairing precomputation on the verification device (ctx device 0)
        DeviceCtx* d = ctx->devs[0];
        if (const char* e = getenv("RAIKO_KZG_PAIRING_LANES")) d->pairing_lanes = atoi(e);
        rk_status st = RK_OK;
        cudaError_t ce = cudaSetDevice(d->dev);
        uint8_t* d_g2 = nullptr;
        if (ce == cudaSuccess) ce = cudaMalloc(&d_g2, 384);
        if (ce == cudaSuccess) ce = cudaMalloc(&d->g2_lines, PAIRING_LINES_BYTES);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_g2, ctx->g2_be.data() + 192, 192, cudaMemcpyHostToDevice);      // [s]G2
        if (ce == cudaSuccess) ce = cudaMemcpy(d_g2 + 192, ctx->g2_be.data(), 192, cudaMemcpyHostToDevice);      // generator
        if (ce == cudaSuccess) {
            launch_k_pairing_precompute(1, 2, 0, 0, d_g2, d_g2 + 192, d->g2_lines);
            ce = cudaDeviceSynchronize();
        }
        cudaFree(d_g2);
        cudaSetDevice(prev);
        if (ce != cudaSuccess) {
            st = fail(RK_ERR_CUDA, "pairing precomputation failed: %s", cudaGetErrorString(ce));
            rk_kzg_ctx_destroy(ctx);
            return st;
        }
    }
    *out = ctx;
    return RK_OK;
}

rk_status rk_kzg_ctx_create(const uint8_t* settings, size_t len, const int* devices, int ndev, rk_kzg_ctx** out) {
    return rk_kzg_ctx_create_ex(settings, len, devices, ndev, 0, out);
}

void rk_kzg_ctx_destroy(rk_kzg_ctx* ctx) {
    if (!ctx) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto* d : ctx->devs) free_device(d);
    cudaSetDevice(prev);
    delete ctx;
}

int rk_kzg_ctx_window_bits(const rk_kzg_ctx* ctx) { return ctx ? ctx->geom.c : 0; }
int rk_kzg_ctx_num_devices(const rk_kzg_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
uint64_t rk_kzg_ctx_table_bytes(const rk_kzg_ctx* ctx) { return ctx ? ctx->devs[0]->table_bytes : 0; }
int rk_kzg_ctx_window_reduced(const rk_kzg_ctx* ctx) {
    if (!ctx) return 0;
    for (auto* d : ctx->devs) if (d->window_reduced) return 1;
    return 0;
}

rk_status rk_kzg_ctx_export_settings(rk_kzg_ctx* ctx, int kind, uint8_t* out, size_t* len) {
    if (!ctx || !out || !len) return fail(RK_ERR_ARG, "null argument");
    const size_t need = kind == 0 ? RAW_LEN : kind == 1 ? BINCODE_LEN : 0;
    if (!need) return fail(RK_ERR_ARG, "kind must be 0 (raw) or 1 (bincode)");
    if (*len < need) { *len = need; return fail(RK_ERR_BAD_LENGTH, "output buffer too small, need %zu", need); }
    DeviceCtx* d = ctx->devs[0];
    DeviceGuard dg;
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    DevBuf buf;
    uint8_t* d_buf = nullptr;
    const size_t g1b = 144ull * NPTS, g2b = 288ull * N_G2, rb = 32ull * 4097;
    CUDA_TRY(buf.alloc(&d_buf, g1b + g2b + 3 * rb + 192 * N_G2));
    uint8_t *d_g1 = d_buf, *d_g2 = d_buf + g1b, *d_r0 = d_g2 + g2b, *d_r1 = d_r0 + rb, *d_r2 = d_r1 + rb, *d_g2in = d_r2 + rb;
    launch_k_setup_to_ref(NPTS / 64, 64, 0, 0, d->g1_aff, NPTS, d_g1);
    // G2: x.c0 x.c1 y.c0 y.c1 (BE) -> 6 Montgomery coordinates with Z = (1, 0)
    std::vector<uint8_t> g2six(288 * N_G2, 0);
    for (int i = 0; i < N_G2; i++) {
        memcpy(g2six.data() + 288 * i, ctx->g2_be.data() + 192 * i, 192);
        g2six[288 * i + 4 * 48 + 47] = 1;
    }
    uint8_t* d_g2be = nullptr;
    CUDA_TRY(buf.alloc(&d_g2be, g2six.size()));
    CUDA_TRY(cudaMemcpy(d_g2be, g2six.data(), g2six.size(), cudaMemcpyHostToDevice));
    launch_k_fp_be_to_ref((6 * N_G2 + 63) / 64, 64, 0, 0, d_g2be, 6 * N_G2, d_g2);
    launch_k_roots_export((4097 + 63) / 64, 64, 0, 0, 0, 4097, d_r0);
    launch_k_roots_export((4097 + 63) / 64, 64, 0, 0, 1, 4097, d_r1);
    launch_k_roots_export((4096 + 63) / 64, 64, 0, 0, 2, 4096, d_r2);
    (void)d_g2in;
    CUDA_TRY(cudaGetLastError());
    std::vector<uint8_t> h(g1b + g2b + 3 * rb);
    CUDA_TRY(cudaMemcpy(h.data(), d_buf, h.size(), cudaMemcpyDeviceToHost));
    const uint8_t *h_g1 = h.data(), *h_g2 = h.data() + g1b, *h_r0 = h_g2 + g2b, *h_r1 = h_r0 + rb, *h_r2 = h_r1 + rb;
    memset(out, 0, need);
    auto put_le64 = [&](size_t off, uint64_t v) { memcpy(out + off, &v, 8); };
    if (kind == 0) {
        out[6] = 0x10;                                   // max_width = 4096, big-endian u64
        memcpy(out + RAW_ROOTS, h_r2, 32 * 4096);
        memcpy(out + RAW_G1, h_g1, g1b);
        memcpy(out + RAW_G2, h_g2, g2b);
    } else {
        put_le64(0, 4096);
        memcpy(out + 8, h_r0 + 32, 32);                  // root_of_unity = w^1
        put_le64(BC_EXPANDED, 4097); memcpy(out + BC_EXPANDED + 8, h_r0, rb);
        put_le64(BC_REVERSE, 4097);  memcpy(out + BC_REVERSE + 8, h_r1, rb);
        put_le64(BC_ROOTS, 4096);    memcpy(out + BC_ROOTS + 8, h_r2, 32 * 4096);
        put_le64(BC_G1, 4096);       memcpy(out + BC_G1 + 8, h_g1, g1b);
        put_le64(BC_G2, N_G2);       memcpy(out + BC_G2 + 8, h_g2, g2b);
        out[BINCODE_LEN - 1] = 0;                        // precomputation: None
    }
    *len = need;
    return RK_OK;
}

rk_status rk_kzg_to_versioned_hash(const uint8_t commitment[48], uint8_t out_hash[32]) {
    if (!commitment || !out_hash) return fail(RK_ERR_ARG, "null argument");
    sha256_short(commitment, 48, out_hash);
    out_hash[0] = RK_VERSIONED_HASH_VERSION_KZG;
    return RK_OK;
}

rk_status rk_commit_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, size_t n, uint8_t* out_commitments,
                          uint8_t* out_versioned_hashes, uint8_t* per_blob_status) {
    if (!out_commitments) return fail(RK_ERR_ARG, "out_commitments is required");
    BatchArgs a{};
    a.mode = MODE_COMMIT; a.blobs = blobs; a.n = n;
    a.out_c = out_commitments; a.out_vh = out_versioned_hashes; a.status = per_blob_status;
    return run_batch(ctx, a);
}

rk_status rk_commit_prove_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, size_t n, uint8_t* out_commitments,
                                uint8_t* out_versioned_hashes, uint8_t* out_x, uint8_t* out_y, uint8_t* out_proofs,
                                uint8_t* per_blob_status) {
    if (!out_commitments) return fail(RK_ERR_ARG, "out_commitments is required");
    BatchArgs a{};
    a.mode = MODE_COMMIT_PROVE; a.blobs = blobs; a.n = n;
    a.out_c = out_commitments; a.out_vh = out_versioned_hashes; a.out_x = out_x; a.out_y = out_y;
    a.out_proof = out_proofs; a.status = per_blob_status;
    return run_batch(ctx, a);
}

rk_status rk_compute_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* zs, size_t n,
                                     uint8_t* out_proofs, uint8_t* out_y, uint8_t* per_blob_status) {
    if (!out_proofs || !zs) return fail(RK_ERR_ARG, "zs and out_proofs are required");
    if (is_device_ptr(zs)) return fail(RK_ERR_ARG, "zs must be a host pointer");
    BatchArgs a{};
    a.mode = MODE_PROOF_AT_Z; a.blobs = blobs; a.zs = zs; a.n = n;
    a.out_proof = out_proofs; a.out_y = out_y; a.status = per_blob_status;
    return run_batch(ctx, a);
}

rk_status rk_compute_blob_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments, size_t n,
                                          uint8_t* out_proofs, uint8_t* per_blob_status) {
    if (!out_proofs || !commitments) return fail(RK_ERR_ARG, "commitments and out_proofs are required");
    if (ctx && is_device_ptr(commitments) && ctx->devs.size() != 1) return fail(RK_ERR_ARG, "device pointers need a single-device context");
    if (ctx) { rk_status pst = check_ptr_device(ctx, commitments, "commitments"); if (pst != RK_OK) return pst; }
    BatchArgs a{};
    a.mode = MODE_BLOB_PROOF; a.blobs = blobs; a.cins = commitments; a.n = n;
    a.out_proof = out_proofs; a.status = per_blob_status;
    return run_batch(ctx, a);
}

// Execute a group of same-mode single-blob requests as one batch and fill in their results.
static void exec_requests(rk_kzg_ctx* ctx, const std::vector<SingleReq*>& batch) {
    const size_t k = batch.size();
    const int mode = batch[0]->mode;
    std::vector<uint8_t> st(k, 0), oc(48 * k), ovh(32 * k), ox(32 * k), oy(32 * k), op(48 * k), aux(32 * k), staged;
    const uint8_t* blobs = batch[0]->blob;            // a lone request is passed through without a staging copy
    if (k > 1) {
        staged.resize(k * (size_t)BLOB_BYTES);
        for (size_t i = 0; i < k; i++) memcpy(staged.data() + i * BLOB_BYTES, batch[i]->blob, BLOB_BYTES);
        blobs = staged.data();
    }
    for (size_t i = 0; i < k; i++) if (batch[i]->aux32) memcpy(aux.data() + 32 * i, batch[i]->aux32, 32);
    BatchArgs a{};
    a.mode = (BatchMode)mode; a.blobs = blobs; a.n = k; a.status = st.data();
    if (mode == MODE_PROOF_AT_Z) a.zs = aux.data(); else a.vhs = aux.data();
    if (mode == MODE_COMMIT) { a.out_c = oc.data(); a.out_vh = ovh.data(); }
    if (mode == MODE_PROOF_AT_Z) { a.out_proof = op.data(); a.out_y = oy.data(); }
    if (mode == MODE_EVAL_ONLY) { a.out_x = ox.data(); a.out_y = oy.data(); }
    if (mode == MODE_POINT_ONLY) { a.out_x = ox.data(); }
    if (mode == MODE_PROVE_VH) { a.out_x = ox.data(); a.out_y = oy.data(); a.out_proof = op.data(); }
    const rk_status rc = run_batch(ctx, a);
    const std::string err = rc != RK_OK ? g_last_error : std::string();
    for (size_t i = 0; i < k; i++) {
        SingleReq* q = batch[i];
        q->rc = rc; q->err = err; q->status = st[i];
        memcpy(q->c, oc.data() + 48 * i, 48); memcpy(q->vh, ovh.data() + 32 * i, 32);
        memcpy(q->x, ox.data() + 32 * i, 32); memcpy(q->y, oy.data() + 32 * i, 32);
        memcpy(q->proof, op.data() + 48 * i, 48);
    }
}

// Run `r` through the batch pipeline, merged with whatever other single-blob requests of the
// same mode are waiting.  Blocks until r is done.
static rk_status submit_single(rk_kzg_ctx* ctx, SingleReq& r) {
    constexpr size_t MAX_MERGE = 64;
    DeviceGuard dg;
    cudaSetDevice(ctx->devs[0]->dev);
    if (is_device_ptr(r.blob)) {                       // device-resident blob: nothing to stage, run alone
        exec_requests(ctx, {&r});
        if (r.rc != RK_OK) g_last_error = r.err;
        return r.rc;
    }
    std::unique_lock<std::mutex> lk(ctx->coal_mu);
    ctx->coal_q.push_back(&r);
    if (ctx->coal_leader) {
        ctx->coal_cv.wait(lk, [&] { return r.done; });
    } else {
        ctx->coal_leader = true;
        while (!ctx->coal_q.empty()) {
            std::vector<SingleReq*> batch;
            const int mode = ctx->coal_q.front()->mode;
            for (auto it = ctx->coal_q.begin(); it != ctx->coal_q.end() && batch.size() < MAX_MERGE;) {
                if ((*it)->mode == mode) { batch.push_back(*it); it = ctx->coal_q.erase(it); } else ++it;
            }
            ctx->coal_batches++;
            ctx->coal_requests += batch.size();
            lk.unlock();
            exec_requests(ctx, batch);
            lk.lock();
            for (SingleReq* q : batch) q->done = true;
            ctx->coal_cv.notify_all();
        }
        ctx->coal_leader = false;
    }
    if (r.rc != RK_OK) g_last_error = r.err;
    return r.rc;
}

static rk_status single_status(uint8_t st) {
    if (st == 0) return RK_OK;
    return fail((rk_status)st, "Failed to deserialize blob to field elements");
}

rk_status rk_blob_to_kzg_commitment(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len, uint8_t out_commitment[48]) {
    if (!ctx || !blob || !out_commitment) return fail(RK_ERR_ARG, "null argument");
    rk_status st = check_blob_len(blob_len);
    if (st != RK_OK) return st;
    SingleReq r;
    r.mode = MODE_COMMIT; r.blob = blob;
    st = submit_single(ctx, r);
    if (st != RK_OK) return st;
    if (r.status) return single_status(r.status);
    memcpy(out_commitment, r.c, 48);
    return RK_OK;
}

rk_status rk_get_evaluation_point(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len, const uint8_t versioned_hash[32],
                                  uint8_t out_x[32]) {
    if (!ctx || !blob || !versioned_hash || !out_x) return fail(RK_ERR_ARG, "null argument");
    rk_status st = check_blob_len(blob_len);
    if (st != RK_OK) return st;
    SingleReq r;
    r.mode = MODE_POINT_ONLY; r.blob = blob; r.aux32 = versioned_hash;
    st = submit_single(ctx, r);          // no deserialisation on this path in the reference either
    if (st != RK_OK) return st;
    memcpy(out_x, r.x, 32);
    return RK_OK;
}

rk_status rk_proof_of_equivalence(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len, const uint8_t versioned_hash[32],
                                  uint8_t out_x[32], uint8_t out_y[32]) {
    if (!ctx || !blob || !versioned_hash || !out_x || !out_y) return fail(RK_ERR_ARG, "null argument");
    rk_status st = check_blob_len(blob_len);
    if (st != RK_OK) return st;
    SingleReq r;
    r.mode = MODE_EVAL_ONLY; r.blob = blob; r.aux32 = versioned_hash;
    st = submit_single(ctx, r);
    if (st != RK_OK) return st;
    if (r.status) return single_status(r.status);
    memcpy(out_x, r.x, 32); memcpy(out_y, r.y, 32);
    return RK_OK;
}

rk_status rk_compute_kzg_proof(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len, const uint8_t z[32],
                               uint8_t out_proof[48], uint8_t out_y[32]) {
    if (!ctx || !blob || !z || !out_proof) return fail(RK_ERR_ARG, "null argument");
    rk_status st = check_blob_len(blob_len);
    if (st != RK_OK) return st;
    SingleReq r;
    r.mode = MODE_PROOF_AT_Z; r.blob = blob; r.aux32 = z;
    st = submit_single(ctx, r);
    if (st != RK_OK) return st;
    if (r.status) return single_status(r.status);
    memcpy(out_proof, r.proof, 48);
    if (out_y) memcpy(out_y, r.y, 32);
    return RK_OK;
}

rk_status rk_calc_kzg_proof(rk_kzg_ctx* ctx, const uint8_t* blob, size_t blob_len, const uint8_t versioned_hash[32],
                            uint8_t out_proof[48]) {
    if (!ctx || !blob || !versioned_hash || !out_proof) return fail(RK_ERR_ARG, "null argument");
    rk_status st = check_blob_len(blob_len);
    if (st != RK_OK) return st;
    SingleReq r;                          // one pass: hash -> challenge -> evaluation -> quotient -> MSM
    r.mode = MODE_PROVE_VH; r.blob = blob; r.aux32 = versioned_hash;
    st = submit_single(ctx, r);
    if (st != RK_OK) return st;
    if (r.status) return single_status(r.status);
    memcpy(out_proof, r.proof, 48);
    return RK_OK;
}

rk_status rk_verify_kzg_proof(rk_kzg_ctx* ctx, const uint8_t commitment[48], const uint8_t z[32], const uint8_t y[32],
                              const uint8_t proof[48], int* out_ok) {
    if (!ctx || !commitment || !z || !y || !proof || !out_ok) return fail(RK_ERR_ARG, "null argument");
    *out_ok = 0;
    DeviceCtx* d = ctx->devs[0];
    DeviceGuard dg;
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    DevBuf buf(&d->arena);
    uint8_t* d_in = nullptr;
    CUDA_TRY(buf.alloc(&d_in, 160));
    uint8_t h[160];
    memcpy(h, commitment, 48); memcpy(h + 48, proof, 48); memcpy(h + 96, z, 32); memcpy(h + 128, y, 32);
    CUDA_TRY(cudaMemcpyAsync(d_in, h, 160, cudaMemcpyHostToDevice, d->s_main));
    return verify_core(ctx, d, buf, 1, d_in, d_in + 48, d_in + 96, d_in + 128, out_ok);
}

rk_status rk_verify_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* commitments, const uint8_t* zs, const uint8_t* ys,
                                    const uint8_t* proofs, size_t n, int* out_ok) {
    if (!ctx || !out_ok) return fail(RK_ERR_ARG, "null argument");
    *out_ok = 0;
    if (n == 0) { *out_ok = 1; return RK_OK; }
    if (!commitments || !zs || !ys || !proofs) return fail(RK_ERR_ARG, "null argument");
    DeviceCtx* d = ctx->devs[0];
    DeviceGuard dg;
    for (const void* p : {(const void*)commitments, (const void*)zs, (const void*)ys, (const void*)proofs}) {
        rk_status pst = check_ptr_device(ctx, p, "an input array");
        if (pst != RK_OK) return pst;
    }
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    // one random-linear-combination transcript (and one two-pairing check) per <= 16384 tuples
    int all = 1;
    for (size_t first = 0; first < n; first += VERIFY_MAX_N) {
        const size_t cnt = std::min(VERIFY_MAX_N, n - first);
        DevBuf buf(&d->arena);
        uint8_t *d_c = nullptr, *d_p = nullptr, *d_z = nullptr, *d_y = nullptr;
        CUDA_TRY(buf.alloc(&d_c, 48 * cnt)); CUDA_TRY(buf.alloc(&d_p, 48 * cnt));
        CUDA_TRY(buf.alloc(&d_z, 32 * cnt)); CUDA_TRY(buf.alloc(&d_y, 32 * cnt));
        cudaStream_t st = d->s_main;
        CUDA_TRY(cudaMemcpyAsync(d_c, commitments + 48 * first, 48 * cnt, cudaMemcpyDefault, st));
        CUDA_TRY(cudaMemcpyAsync(d_p, proofs + 48 * first, 48 * cnt, cudaMemcpyDefault, st));
        CUDA_TRY(cudaMemcpyAsync(d_z, zs + 32 * first, 32 * cnt, cudaMemcpyDefault, st));
        CUDA_TRY(cudaMemcpyAsync(d_y, ys + 32 * first, 32 * cnt, cudaMemcpyDefault, st));
        int ok = 0;
        rk_status rc = verify_core(ctx, d, buf, (int)cnt, d_c, d_p, d_z, d_y, &ok);
        if (rc != RK_OK) return rc;
        all &= ok;
    }
    *out_ok = all;
    return RK_OK;
}

rk_status rk_verify_blob_kzg_proof_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments,
                                         const uint8_t* proofs, size_t n, int* out_ok) {
    if (!ctx || !out_ok) return fail(RK_ERR_ARG, "null argument");
    *out_ok = 0;
    if (n == 0) { *out_ok = 1; return RK_OK; }
    if (!blobs || !commitments || !proofs) return fail(RK_ERR_ARG, "null argument");
    // one transcript per <= 16384 blobs; larger inputs are verified as consecutive sub-batches
    int all = 1;
    for (size_t first = 0; first < n; first += VERIFY_MAX_N) {
        const size_t cnt = std::min(VERIFY_MAX_N, n - first);
        int ok = 0;
        rk_status st = verify_blob_batch_device(ctx, blobs + first * BLOB_BYTES, commitments + 48 * first, proofs + 48 * first, cnt, &ok);
        if (st != RK_OK) return st;
        all &= ok;
    }
    *out_ok = all;
    return RK_OK;
}

rk_status rk_decode_blob_data_batch(rk_kzg_ctx* ctx, const uint8_t* blobs, size_t n, uint8_t* out, uint32_t* out_len) {
    if (!ctx || !out || !out_len) return fail(RK_ERR_ARG, "null argument");
    if (n == 0) return RK_OK;
    if (!blobs) return fail(RK_ERR_ARG, "null blobs pointer");
    DeviceCtx* d = ctx->devs[0];
    DeviceGuard dg;
    for (const void* p : {(const void*)blobs, (const void*)out, (const void*)out_len}) {
        rk_status pst = check_ptr_device(ctx, p, "a buffer");
        if (pst != RK_OK) return pst;
    }
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    const bool blobs_dev = is_device_ptr(blobs), out_dev = is_device_ptr(out);
    if (out_dev != is_device_ptr(out_len)) return fail(RK_ERR_ARG, "out and out_len must both be host or both be device pointers");
    if (blobs_dev && ((uintptr_t)blobs & 15)) return fail(RK_ERR_ARG, "device blob pointer must be 16-byte aligned");
    const size_t chunk = (size_t)d->chunk;
    cudaStream_t st = d->s_main;
    DevBuf buf;
    uint32_t* d_len = nullptr;
    CUDA_TRY(buf.alloc(&d_len, chunk));
    for (size_t first = 0; first < n; first += chunk) {
        ChunkSlot& s = d->slot[0];
        const int cnt = (int)std::min(chunk, n - first);
        const uint8_t* d_blobs;
        if (blobs_dev) {
            d_blobs = blobs + first * BLOB_BYTES;
        } else {
            CUDA_TRY(cudaMemcpyAsync(s.d_blobs, blobs + first * BLOB_BYTES, (size_t)cnt * BLOB_BYTES, cudaMemcpyHostToDevice, st));
            d->stats.h2d_bytes += (uint64_t)cnt * BLOB_BYTES;
            d_blobs = s.d_blobs;
        }
        uint8_t* d_o = out_dev ? out + first * BLOBDATA_STRIDE : s.d_q;          // d_q: chunk * 131072 >= chunk * stride
        uint32_t* d_l = out_dev ? out_len + first : d_len;
        launch_k_decode_blob_data(cnt, 256, 0, st, d_blobs, cnt, d_o, d_l);
        d->stats.total_launches++;
        CUDA_TRY(cudaGetLastError());
        if (!out_dev) {
            CUDA_TRY(cudaMemcpyAsync(out + first * BLOBDATA_STRIDE, d_o, (size_t)cnt * BLOBDATA_STRIDE, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(out_len + first, d_l, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, st));
            d->stats.d2h_bytes += (uint64_t)cnt * (BLOBDATA_STRIDE + 4);
        }
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return RK_OK;
}

rk_status rk_synth_blobs(rk_kzg_ctx* ctx, uint64_t seed, uint32_t first_blob, size_t n, uint8_t* out) {
    if (!ctx || !out) return fail(RK_ERR_ARG, "null argument");
    if (n == 0) return RK_OK;
    if (n > 0xffffffffull - first_blob) return fail(RK_ERR_ARG, "blob index exceeds 32 bits");
    DeviceCtx* d = ctx->devs[0];
    DeviceGuard dg;
    rk_status pst = check_ptr_device(ctx, out, "out");
    if (pst != RK_OK) return pst;
    std::lock_guard<std::mutex> lock(d->mu);
    CUDA_TRY(cudaSetDevice(d->dev));
    const bool out_dev = is_device_ptr(out);
    const size_t chunk = (size_t)d->chunk;
    for (size_t first = 0; first < n; first += chunk) {
        const size_t cnt = std::min(chunk, n - first);
        uint8_t* dst = out_dev ? out + first * BLOB_BYTES : d->slot[0].d_q;
        launch_k_synth_blobs((unsigned)(cnt * NPTS / 256), 256, 0, d->s_main, seed, first_blob + (uint32_t)first, (uint32_t)cnt, dst);
        d->stats.total_launches++;
        CUDA_TRY(cudaGetLastError());
        if (!out_dev) CUDA_TRY(cudaMemcpyAsync(out + first * BLOB_BYTES, dst, cnt * BLOB_BYTES, cudaMemcpyDeviceToHost, d->s_main));
        CUDA_TRY(cudaStreamSynchronize(d->s_main));
    }
    return RK_OK;
}

void rk_kzg_stats_enable(rk_kzg_ctx* ctx, int enable) {
    if (ctx) for (auto* d : ctx->devs) d->stats_on = enable != 0;
}
void rk_kzg_stats_reset(rk_kzg_ctx* ctx) {
    if (ctx) for (auto* d : ctx->devs) { std::lock_guard<std::mutex> l(d->mu); d->stats = rk_kzg_stats{}; }
}
void rk_kzg_stats_get(rk_kzg_ctx* ctx, rk_kzg_stats* out) {
    if (!ctx || !out) return;
    *out = rk_kzg_stats{};
    for (auto* d : ctx->devs) {
        std::lock_guard<std::mutex> l(d->mu);
        out->msm_ms += d->stats.msm_ms; out->fr_ms += d->stats.fr_ms; out->sha_ms += d->stats.sha_ms;
        out->finalize_ms += d->stats.finalize_ms; out->msm_launches += d->stats.msm_launches;
        out->total_launches += d->stats.total_launches; out->h2d_bytes += d->stats.h2d_bytes;
        out->d2h_bytes += d->stats.d2h_bytes; out->msm_point_adds += d->stats.msm_point_adds;
        out->msm_affine_launches += d->stats.msm_affine_launches; out->msm_affine_point_adds += d->stats.msm_affine_point_adds;
    }
}

rk_status rk_measure_imad_peak(int device, double* out_macs_per_sec, double* out_sm_clock_mhz) {
    if (!out_macs_per_sec) return fail(RK_ERR_ARG, "null argument");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return fail(RK_ERR_CUDA, "no such CUDA device %d", device);
    DeviceGuard dg;
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, iters = 8192;
    uint64_t* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(uint64_t) * blocks * 256));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        launch_k_imad_peak(blocks, 256, 0, 0, d_out, 3, iters);
        cudaEventRecord(e1);
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    *out_macs_per_sec = (double)blocks * 256 * iters * 16 * 8 / (best * 1e-3);
    if (out_sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *out_sm_clock_mhz = khz / 1000.0;
    }
    return RK_OK;
}

}  // extern "C"
