// Host engine + kernels of the B200 A*PA2 path, exported through the C-ABI in include/astarpa.h and
// include/astarpa_b200.h.  There is NO CPU fallback: every entry point fails loudly without a usable device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/astarpa.h"
#include "../../include/astarpa_b200.h"
#include "apa_gcsh.cuh"
#include "apa_trace.cuh"

using namespace apa;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
static int set_err(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            return set_err(APA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

// ------------------------------------------------------------------------------------------------ kernels
struct BatchDev {
    uint64_t n_pairs;
    const uint8_t* a_all;
    const uint8_t* b_all;
    const int64_t* a_off;
    const int64_t* b_off;
    const int64_t* bp_off;  // per pair offset into bprof, in 32-row half-words (two padding half-words per pair)
    uint2* bprof;
    const int64_t* ap_off;  // the same for a
    uint2* aprof;
    int32_t* status;  // per pair
    int32_t* cost;    // per pair
    int64_t* cig_off;
    int64_t* cig_len;
    const uint32_t* order;  // work order (largest first)
    uint32_t n_order;
    unsigned long long* queue;  // work-queue head
    uint8_t* arena;             // n_slots * arena_size
    uint32_t arena_size;
    char* pool;
    unsigned long long* pool_cursor;
    unsigned long long pool_cap;
    unsigned long long* stats;  // [0] word_steps [1] computed_cells [2] passes [3] fill_blocks [4] dt_blocks [5..12] phase cycles
    int preset;
    int trace;
    const volatile uint32_t* ready;  // streaming upload: number of pairs (in work order) whose bases are in HBM; nullptr = all
    int32_t* dbg;  // band log of the (single) pair, or nullptr
    uint32_t dbg_cap;
    uint32_t* dbg_n;
};

// K0 (fused into the align kernel): negated bit planes per 32 bases (BitProfile::build for b,
// pa-bitpacking/src/profile.rs:124-131; the same packing of a feeds the diagonal extensions) and input validation
// (the reference panics on bytes outside ACGT, profile.rs:113). Each lane packs one 32-base group per iteration.
__device__ bool dev_pack_planes(const uint8_t* __restrict__ seq, int64_t len, uint2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t nhw = ((len + 63) / 64) * 2 + 2;  // + two zero half-words of padding
    bool bad = false;
    for (int64_t hw = lane; hw < nhw; hw += 32) {
        uint32_t b0 = 0, b1 = 0;
        const int64_t j0 = hw * 32;
        if (j0 + 32 <= len && ((reinterpret_cast<uintptr_t>(seq + j0) & 3) == 0)) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(seq + j0);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint32_t x = w[q];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    uint32_t c = (x >> (8 * t)) & 0xffu;
                    bad |= !is_acgt(c);
                    uint32_t r = rank_acgt(c);
                    b0 |= ((r & 1u) ^ 1u) << (4 * q + t);
                    b1 |= ((r >> 1) ^ 1u) << (4 * q + t);
                }
            }
        } else {
            for (int t = 0; t < 32; t++) {
                int64_t j = j0 + t;
                if (j < len) {
                    uint32_t c = seq[j];
                    bad |= !is_acgt(c);
                    uint32_t r = rank_acgt(c);
                    b0 |= ((r & 1u) ^ 1u) << t;
                    b1 |= ((r >> 1) ^ 1u) << t;
                }
            }
        }
        out[hw] = make_uint2(b0, b1);
    }
    bad = __any_sync(FULL, bad);
    __syncwarp();
    return !bad;
}

constexpr int WARPS_PER_CTA = 4;

// K1+K3 fused per pair: a persistent warp pulls pairs from the device work queue and runs the band-doubling
// search, the traceback and the CIGAR text emission for each.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 8) apa_align_kernel(BatchDev bd) {
    __shared__ WarpSmem smem[WARPS_PER_CTA];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    WarpSmem& sm = smem[wib];
    const uint32_t slot = blockIdx.x * WARPS_PER_CTA + wib;
    uint8_t* arena = bd.arena + (size_t)slot * bd.arena_size;
    unsigned long long acc_steps = 0, acc_cells = 0, acc_pass = 0, acc_fill = 0, acc_dt = 0;
    long long acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(bd.queue, 1ull);
        q = __shfl_sync(FULL, q, 0);
        if (q >= bd.n_order) break;
        if (bd.ready) {  // streaming upload: wait until this pair's bases have landed in HBM
            if (lane == 0) {
                while (*bd.ready <= (uint32_t)q) __nanosleep(500);
            }
            __syncwarp();
        }
        const uint32_t p = bd.order[q];

        PairCtx cx;
        cx.n = (I)(bd.a_off[p + 1] - bd.a_off[p]);
        cx.m = (I)(bd.b_off[p + 1] - bd.b_off[p]);
        cx.a = bd.a_all + bd.a_off[p];
        cx.b = bd.b_all + bd.b_off[p];
        cx.bprof = bd.bprof + bd.bp_off[p];
        cx.aprof = bd.aprof + bd.ap_off[p];
        {
            bool ok_a = dev_pack_planes(cx.a, cx.n, bd.aprof + bd.ap_off[p]);
            bool ok_b = dev_pack_planes(cx.b, cx.m, bd.bprof + bd.bp_off[p]);
            if (!(ok_a && ok_b)) {
                if (lane == 0) bd.status[p] = ST_BAD_INPUT;
                continue;
            }
        }
        cx.arena = arena;
        cx.arena_size = bd.arena_size;
        cx.nblk = (cx.n + BLOCK_W - 1) / BLOCK_W;
        cx.nblk_alloc = 0;
        cx.meta = (BlkMeta*)arena;
        uint32_t meta_bytes = ((uint32_t)(cx.nblk + 1) * (uint32_t)sizeof(BlkMeta) + 15u) & ~15u;
        cx.v_base = meta_bytes;
        cx.v_top = meta_bytes;
        cx.hi_bot = bd.arena_size;
        cx.status = ST_PENDING;
        cx.word_steps = cx.computed_cells = 0;
        cx.passes = 0;
        cx.fill_blocks = cx.dt_blocks = 0;
        for (int t = 0; t < 8; t++) cx.tphase[t] = 0;
        cx.dbg = bd.dbg;
        cx.dbg_cap = bd.dbg_cap;
        cx.dbg_n = 0;
        if (meta_bytes + 4096u > bd.arena_size) cx.status = ST_OVERFLOW;

        Cost cost = -1;
        long long cig_off = -1, cig_len = 0;
        if (cx.status == ST_PENDING) {
            if (bd.preset == APA_PRESET_SIMPLE) {
                GapH hh{cx.n, cx.m};
                Cost h0 = hh.h(0, 0);
                long long t0 = APA_TIC();
                cost = dev_band_doubling(cx, sm, hh, h0);
                APA_TOC(cx.tphase[2], t0);
                if (cx.status == ST_PENDING && h0 > cost) cx.status = ST_ASSERT;  // lib.rs:173
            } else {
                GcshH hh;
                long long t0 = APA_TIC();
                bool built = gcsh_build(cx, hh);
                APA_TOC(cx.tphase[0], t0);
                if (built) {
                    Cost h0 = hh.h(0, 0);
                    t0 = APA_TIC();
                    cost = dev_band_doubling(cx, sm, hh, h0);
                    APA_TOC(cx.tphase[2], t0);
                    cx.tphase[6] += hh.t_h;
                    if (cx.status == ST_PENDING && h0 > cost) cx.status = ST_ASSERT;  // lib.rs:173
                }
            }
        }
        if (cx.status == ST_PENDING && bd.trace) {
            CigarWriter cw;
            cw.arena = arena;
            cw.arena_size = bd.arena_size;
            cw.count = 0;
            cw.pend_cnt = 0;
            cw.pend_op = 0;
            long long t0 = APA_TIC();
            bool traced = dev_trace(cx, sm, cw, cost);
            APA_TOC(cx.tphase[3], t0);
            if (traced) {
                t0 = APA_TIC();
                cig_off = emit_cigar_text(cw, bd.pool, bd.pool_cursor, bd.pool_cap, &cig_len);
                APA_TOC(cx.tphase[5], t0);
                if (cig_off < 0) cx.status = ST_OVERFLOW;
            }
        }
        if (lane == 0) {
            bd.status[p] = cx.status == ST_PENDING ? ST_DONE : cx.status;
            bd.cost[p] = cost;
            bd.cig_off[p] = cig_off;
            bd.cig_len[p] = cig_len;
        }
        if (bd.dbg_n && lane == 0) *bd.dbg_n = cx.dbg_n;
        acc_steps += cx.word_steps;
        acc_cells += cx.computed_cells;
        acc_pass += cx.passes;
        acc_fill += cx.fill_blocks;
        acc_dt += cx.dt_blocks;
        for (int t = 0; t < 8; t++) acc_t[t] += cx.tphase[t];
    }
    if (lane == 0) {
        atomicAdd(&bd.stats[0], acc_steps);
        atomicAdd(&bd.stats[1], acc_cells);
        atomicAdd(&bd.stats[2], acc_pass);
        atomicAdd(&bd.stats[3], acc_fill);
        atomicAdd(&bd.stats[4], acc_dt);
        for (int t = 0; t < 8; t++) atomicAdd(&bd.stats[5 + t], (unsigned long long)acc_t[t]);
    }
}

// Stand-alone block-DP rectangle (apa_block_compute): one warp, arbitrary top deltas are not needed by the hot
// path (HMode::None only), so h_in must be all +1; h_out is reconstructed column by column for the KAT.
__global__ void apa_block_kernel(const uint8_t* a, int na, const uint2* bprof, int nhw, uint2* v, int32_t* cum, uint2* fillvals) {
    __shared__ WarpSmem sm;
    const int lane = threadIdx.x & 31;
    BlkView prev;
    prev.js = 0;
    prev.je = nhw * 32;
    prev.top_val = 0;
    prev.bot_val = 0;
    prev.v = v;
    prev.cum = nullptr;
    prev.ones = 0;
    unsigned long long ws = 0;
    // the rectangle may be wider than one block: sweep it in 256-column slabs, each slab's right column feeding the next
    for (int c0 = 0; c0 < na; c0 += BLOCK_W) {
        int nc = min(BLOCK_W, na - c0);
        stage_amask(sm, a, c0, nc, lane);
        block_dp<true>(sm, bprof, prev, nc, 0, nhw * 32, v, cum, 0, fillvals + (size_t)c0 * nhw, ws);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ host side
struct apa_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // streaming uploads overlap the persistent kernel
    uint32_t* d_ready = nullptr;
    uint32_t* h_ready = nullptr;  // pinned: cumulative pair counts per upload chunk
    cudaEvent_t ev[6] = {};
    unsigned long long* d_queue = nullptr;   // [0] queue head, [1] pool cursor, [2..] stats
    uint8_t* d_arena = nullptr;
    size_t arena_total = 0;
};

struct apa_batch {
    uint64_t n_pairs = 0;
    std::vector<int64_t> a_off, b_off, bp_off, ap_off;
    uint64_t total_a = 0, total_b = 0, total_hw = 0, total_hw_a = 0;
    I max_n = 0, max_m = 0;
    uint8_t *d_a = nullptr, *d_b = nullptr;
    int64_t *d_a_off = nullptr, *d_b_off = nullptr, *d_bp_off = nullptr, *d_ap_off = nullptr;
    uint2 *d_bprof = nullptr, *d_aprof = nullptr;
    int32_t *d_status = nullptr, *d_cost = nullptr;
    int64_t *d_cig_off = nullptr, *d_cig_len = nullptr;
    uint32_t* d_order = nullptr;
    char* d_pool = nullptr;
    uint64_t pool_cap = 0;
    // streaming upload (apa_align_batch): bases are copied chunk by chunk while the kernel already runs
    const uint8_t *h_a = nullptr, *h_b = nullptr;
    std::vector<uint32_t> chunk_end;  // pairs (in work order) available after each chunk
    std::vector<uint32_t> chunk_pair_end;  // pair index (exclusive) of each chunk: chunks are contiguous index ranges
    bool ran = false;
    int trace = 0;
    apa_batch_stats stats{};
    std::vector<int32_t> h_status;
    uint64_t pool_used = 0;
    int32_t* d_dbg = nullptr;  // set by apa_debug_band_log
    uint32_t dbg_cap = 0;
    uint32_t* d_dbg_n = nullptr;
};

extern "C" const char* apa_last_error(void) { return g_last_error.c_str(); }

extern "C" int apa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int apa_engine_create(int device, apa_engine** out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(APA_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                                              " (this library has no CPU fallback)");
    if (device < 0 || device >= n) return set_err(APA_ERR_NO_DEVICE, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return set_err(APA_ERR_NO_DEVICE, "built for sm_100a; device is older");
    apa_engine* eng = new apa_engine();
    eng->device = device;
    eng->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&eng->d_ready, 64));
    CUDA_TRY(cudaHostAlloc((void**)&eng->h_ready, 256 * sizeof(uint32_t), cudaHostAllocDefault));
    for (auto& ev : eng->ev) CUDA_TRY(cudaEventCreate(&ev));
    CUDA_TRY(cudaMalloc(&eng->d_queue, 16 * sizeof(unsigned long long)));
    *out = eng;
    return APA_OK;
}

extern "C" void apa_engine_destroy(apa_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->d_arena) cudaFree(e->d_arena);
    if (e->d_queue) cudaFree(e->d_queue);
    for (auto& ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->d_ready) cudaFree(e->d_ready);
    if (e->h_ready) cudaFreeHost(e->h_ready);
    delete e;
}

extern "C" void apa_free(void* p) { free(p); }

// Pinned (page-locked) host memory for the end-to-end path.
extern "C" void* apa_pinned_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        set_err(APA_ERR_CUDA, "cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
extern "C" void apa_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" void apa_batch_free(apa_engine* e, apa_batch* b) {
    if (!b) return;
    if (e) cudaSetDevice(e->device);
    cudaFree(b->d_a);
    cudaFree(b->d_b);
    cudaFree(b->d_a_off);
    cudaFree(b->d_b_off);
    cudaFree(b->d_bp_off);
    cudaFree(b->d_bprof);
    cudaFree(b->d_aprof);
    cudaFree(b->d_ap_off);
    cudaFree(b->d_status);
    cudaFree(b->d_cost);
    cudaFree(b->d_cig_off);
    cudaFree(b->d_cig_len);
    cudaFree(b->d_order);
    cudaFree(b->d_pool);
    delete b;
}

static int batch_prepare(apa_engine* e, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off, const uint8_t* b_all,
                         const int64_t* b_off, bool defer_data, apa_batch** out) {
    *out = nullptr;
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    apa_batch* b = new apa_batch();
    b->n_pairs = n_pairs;
    b->a_off.assign(a_off, a_off + n_pairs + 1);
    b->b_off.assign(b_off, b_off + n_pairs + 1);
    b->bp_off.resize(n_pairs + 1);
    b->ap_off.resize(n_pairs + 1);
    uint64_t hw = 0, hwa = 0;
    for (uint64_t p = 0; p < n_pairs; p++) {
        int64_t n = a_off[p + 1] - a_off[p], m = b_off[p + 1] - b_off[p];
        if (n < 0 || m < 0 || n >= (1ll << 31) - 1024 || m >= (1ll << 31) - 1024) {
            delete b;
            return set_err(APA_ERR_TOO_LARGE, "sequence length must be < 2^31 (I = i32)");
        }
        b->max_n = std::max<I>(b->max_n, (I)n);
        b->max_m = std::max<I>(b->max_m, (I)m);
        b->bp_off[p] = (int64_t)hw;
        hw += (uint64_t)((m + 63) / 64) * 2 + 2;
        b->ap_off[p] = (int64_t)hwa;
        hwa += (uint64_t)((n + 63) / 64) * 2 + 2;
    }
    b->bp_off[n_pairs] = (int64_t)hw;
    b->ap_off[n_pairs] = (int64_t)hwa;
    b->total_hw = hw;
    b->total_hw_a = hwa;
    b->total_a = (uint64_t)(a_off[n_pairs] - a_off[0]);
    b->total_b = (uint64_t)(b_off[n_pairs] - b_off[0]);
    // Rebase offsets to the start of the copied ranges.
    std::vector<int64_t> ao(b->a_off), bo(b->b_off);
    for (auto& x : ao) x -= a_off[0];
    for (auto& x : bo) x -= b_off[0];
    b->a_off = ao;
    b->b_off = bo;
    // Work order: largest estimated work first. With a streaming upload the batch is cut into chunks of
    // contiguous pairs (about 32 MB of bases each) that become available one after the other; pairs are sorted inside
    // each chunk only.
    std::vector<uint32_t> order(n_pairs);
    std::iota(order.begin(), order.end(), 0u);
    auto by_size = [&](uint32_t x, uint32_t y) {
        return (ao[x + 1] - ao[x]) + (bo[x + 1] - bo[x]) > (ao[y + 1] - ao[y]) + (bo[y + 1] - bo[y]);
    };
    if (defer_data && n_pairs) {
        const uint64_t total = b->total_a + b->total_b;
        const uint64_t n_chunks = std::min<uint64_t>(200, std::max<uint64_t>(1, total / (32ull << 20)));
        const uint64_t per = (total + n_chunks - 1) / n_chunks;
        uint64_t acc = 0, start = 0;
        for (uint64_t p = 0; p < n_pairs; p++) {
            acc += (uint64_t)(ao[p + 1] - ao[p]) + (uint64_t)(bo[p + 1] - bo[p]);
            if (acc >= per || p + 1 == n_pairs) {
                std::stable_sort(order.begin() + start, order.begin() + p + 1, by_size);
                b->chunk_pair_end.push_back((uint32_t)(p + 1));
                acc = 0;
                start = p + 1;
            }
        }
        b->h_a = a_all + a_off[0];
        b->h_b = b_all + b_off[0];
    } else {
        std::stable_sort(order.begin(), order.end(), by_size);
    }

    cudaStream_t st = e->stream;
    CUDA_TRY(cudaEventRecord(e->ev[0], st));
    CUDA_TRY(cudaMalloc(&b->d_a, std::max<uint64_t>(b->total_a, 16) + 64));
    CUDA_TRY(cudaMalloc(&b->d_b, std::max<uint64_t>(b->total_b, 16) + 64));
    CUDA_TRY(cudaMalloc(&b->d_a_off, (n_pairs + 1) * 8));
    CUDA_TRY(cudaMalloc(&b->d_b_off, (n_pairs + 1) * 8));
    CUDA_TRY(cudaMalloc(&b->d_bp_off, (n_pairs + 1) * 8));
    CUDA_TRY(cudaMalloc(&b->d_bprof, std::max<uint64_t>(hw, 2) * 8));
    CUDA_TRY(cudaMalloc(&b->d_aprof, std::max<uint64_t>(hwa, 2) * 8));
    CUDA_TRY(cudaMalloc(&b->d_ap_off, (n_pairs + 1) * 8));
    CUDA_TRY(cudaMalloc(&b->d_status, std::max<uint64_t>(n_pairs, 1) * 4));
    CUDA_TRY(cudaMalloc(&b->d_cost, std::max<uint64_t>(n_pairs, 1) * 4));
    CUDA_TRY(cudaMalloc(&b->d_cig_off, std::max<uint64_t>(n_pairs, 1) * 8));
    CUDA_TRY(cudaMalloc(&b->d_cig_len, std::max<uint64_t>(n_pairs, 1) * 8));
    CUDA_TRY(cudaMalloc(&b->d_order, std::max<uint64_t>(n_pairs, 1) * 8));  // [0,n): work order, [n,2n): retry list
    if (!defer_data) {
        if (b->total_a) CUDA_TRY(cudaMemcpyAsync(b->d_a, a_all + a_off[0], b->total_a, cudaMemcpyHostToDevice, st));
        if (b->total_b) CUDA_TRY(cudaMemcpyAsync(b->d_b, b_all + b_off[0], b->total_b, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_a_off, ao.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_b_off, bo.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_bp_off, b->bp_off.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_ap_off, b->ap_off.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_pairs) CUDA_TRY(cudaMemcpyAsync(b->d_order, order.data(), n_pairs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(e->ev[1], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
    b->stats.h2d_ms = ms;
    b->stats.h2d_bytes = b->total_a + b->total_b + 4 * (n_pairs + 1) * 8 + n_pairs * 4;
    *out = b;
    return APA_OK;
}

extern "C" int apa_batch_upload(apa_engine* e, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off, const uint8_t* b_all,
                                const int64_t* b_off, apa_batch** out) {
    return batch_prepare(e, n_pairs, a_all, a_off, b_all, b_off, false, out);
}

static uint32_t estimate_arena(const apa_batch* b, int preset, int trace) {
    // meta + V columns of one pass + traceback scratch + CIGAR elements. Deliberately modest: pairs that do
    // not fit are re-run with a larger arena (ST_OVERFLOW).
    uint64_t nblk = (uint64_t)(b->max_n + BLOCK_W - 1) / BLOCK_W + 1;
    uint64_t meta = nblk * sizeof(BlkMeta);
    uint64_t band_rows = preset == APA_PRESET_SIMPLE ? std::min<uint64_t>((uint64_t)b->max_m + 64, std::max<uint64_t>(2048, (uint64_t)b->max_n / 8))
                                                     : std::min<uint64_t>((uint64_t)b->max_m + 64, 2048);
    uint64_t vcols = nblk * (band_rows / 32 * 12 + 16);
    uint64_t tr = trace ? (DT_CACHE_ELEMS * 8 + 256 * (band_rows / 32) * 8 / 4 + (uint64_t)(b->max_n + b->max_m) / 4 * 4 + 65536) : 0;
    uint64_t heur = preset == APA_PRESET_FULL ? 24ull * (uint64_t)b->max_n + 65536 : 0;  // k-mer table, matches, contours
    uint64_t s = meta + vcols + tr + heur + 16384;
    s = (s + 1023) & ~1023ull;
    return (uint32_t)std::min<uint64_t>(s, 0xF0000000ull);
}

static int batch_run(apa_engine* e, apa_batch* b, int preset, int trace, bool stream_data) {
    if (!e || !b) return set_err(APA_ERR_NO_DEVICE, "null engine/batch");
    if (preset != APA_PRESET_SIMPLE && preset != APA_PRESET_FULL) return set_err(APA_ERR_BAD_INPUT, "unknown preset");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    b->trace = trace;
    b->stats.kernel_launches = 0;
    b->stats.retries = 0;
    if (b->n_pairs == 0) {
        b->ran = true;
        return APA_OK;
    }
    // CIGAR text pool: text length <= |a| + |b| per pair (+ NUL).
    uint64_t pool_need = trace ? (b->total_a + b->total_b + b->n_pairs + 64) : 16;
    if (b->pool_cap < pool_need) {
        cudaFree(b->d_pool);
        b->d_pool = nullptr;
        CUDA_TRY(cudaMalloc(&b->d_pool, pool_need));
        b->pool_cap = pool_need;
    }
    BatchDev bd{};
    bd.n_pairs = b->n_pairs;
    bd.a_all = b->d_a;
    bd.b_all = b->d_b;
    bd.a_off = b->d_a_off;
    bd.b_off = b->d_b_off;
    bd.bp_off = b->d_bp_off;
    bd.bprof = b->d_bprof;
    bd.ap_off = b->d_ap_off;
    bd.aprof = b->d_aprof;
    bd.status = b->d_status;
    bd.cost = b->d_cost;
    bd.cig_off = b->d_cig_off;
    bd.cig_len = b->d_cig_len;
    bd.order = b->d_order;
    bd.n_order = (uint32_t)b->n_pairs;
    bd.queue = e->d_queue;
    bd.pool = b->d_pool;
    bd.pool_cursor = e->d_queue + 1;
    bd.pool_cap = b->pool_cap;
    bd.stats = e->d_queue + 2;
    bd.preset = preset;
    bd.trace = trace;
    bd.dbg = b->d_dbg;
    bd.dbg_cap = b->dbg_cap;
    bd.dbg_n = b->d_dbg_n;

    CUDA_TRY(cudaEventRecord(e->ev[2], st));
    CUDA_TRY(cudaMemsetAsync(e->d_queue, 0, 16 * sizeof(unsigned long long), st));
    CUDA_TRY(cudaMemsetAsync(b->d_status, 0, b->n_pairs * 4, st));
    uint32_t arena_size = estimate_arena(b, preset, trace);
    std::vector<uint32_t> pending;  // empty = all pairs in the uploaded order
    b->h_status.assign(b->n_pairs, 0);
    for (int attempt = 0; attempt < 8; attempt++) {
        int ctas_per_sm = 8;
        uint64_t want_slots = (uint64_t)e->sm_count * ctas_per_sm * WARPS_PER_CTA;
        uint64_t n_work = attempt == 0 ? b->n_pairs : pending.size();
        uint64_t slots = std::min<uint64_t>(want_slots, ((n_work + WARPS_PER_CTA - 1) / WARPS_PER_CTA) * WARPS_PER_CTA);
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        uint64_t budget = (uint64_t)free_b + e->arena_total;
        budget = budget > (4ull << 30) ? budget - (2ull << 30) : budget / 2;
        while (slots > WARPS_PER_CTA && slots * (uint64_t)arena_size > budget) slots = (slots / 2 / WARPS_PER_CTA) * WARPS_PER_CTA;
        if (slots * (uint64_t)arena_size > budget) return set_err(APA_ERR_TOO_LARGE, "scratch arena exceeds device memory");
        size_t need = (size_t)slots * arena_size;
        if (e->arena_total < need) {
            if (e->d_arena) CUDA_TRY(cudaFree(e->d_arena));
            e->d_arena = nullptr;
            e->arena_total = 0;
            CUDA_TRY(cudaMalloc(&e->d_arena, need));
            e->arena_total = need;
        }
        bd.arena = e->d_arena;
        bd.arena_size = arena_size;
        bd.n_order = (uint32_t)n_work;
        if (attempt > 0) {
            CUDA_TRY(cudaMemcpyAsync(b->d_order + b->n_pairs, pending.data(), pending.size() * 4, cudaMemcpyHostToDevice, st));
            bd.order = b->d_order + b->n_pairs;
            CUDA_TRY(cudaMemsetAsync(e->d_queue, 0, sizeof(unsigned long long), st));
        }
        const bool streaming = stream_data && attempt == 0 && !b->chunk_pair_end.empty();
        bd.ready = streaming ? e->d_ready : nullptr;
        if (streaming) {
            CUDA_TRY(cudaMemsetAsync(e->d_ready, 0, 4, st));
            CUDA_TRY(cudaStreamSynchronize(st));  // offsets, order, zeroed queue are in place before anything overlaps
        }
        apa_align_kernel<<<(unsigned)(slots / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(bd);
        b->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        if (streaming) {
            // H2D of the bases, chunk by chunk on the copy stream, while the persistent kernel already consumes them.
            uint32_t p0 = 0;
            for (size_t c = 0; c < b->chunk_pair_end.size(); c++) {
                const uint32_t p1 = b->chunk_pair_end[c];
                const int64_t a0 = b->a_off[p0], a1 = b->a_off[p1], b0 = b->b_off[p0], b1 = b->b_off[p1];
                if (a1 > a0) CUDA_TRY(cudaMemcpyAsync(b->d_a + a0, b->h_a + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, e->copy_stream));
                if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(b->d_b + b0, b->h_b + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, e->copy_stream));
                e->h_ready[c] = p1;
                CUDA_TRY(cudaMemcpyAsync(e->d_ready, &e->h_ready[c], 4, cudaMemcpyHostToDevice, e->copy_stream));
                p0 = p1;
            }
            CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
        }
        CUDA_TRY(cudaMemcpyAsync(b->h_status.data(), b->d_status, b->n_pairs * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        pending.clear();
        for (uint64_t p = 0; p < b->n_pairs; p++)
            if (b->h_status[p] == ST_OVERFLOW) pending.push_back((uint32_t)p);
        if (pending.empty()) break;
        b->stats.retries += pending.size();
        for (uint32_t p : pending) b->h_status[p] = ST_PENDING;
        // reset the status of the overflowed pairs and grow the arena
        std::vector<int32_t> zero(1, 0);
        for (uint32_t p : pending) CUDA_TRY(cudaMemcpyAsync(b->d_status + p, zero.data(), 4, cudaMemcpyHostToDevice, st));
        uint64_t grown = (uint64_t)arena_size * 4;
        if (grown > 0xF0000000ull) {
            if (arena_size >= 0xF0000000u) return set_err(APA_ERR_TOO_LARGE, "pair does not fit the largest scratch arena");
            grown = 0xF0000000ull;
        }
        arena_size = (uint32_t)grown;
    }
    if (!pending.empty()) return set_err(APA_ERR_TOO_LARGE, "pairs still overflow after 8 arena enlargements");
    CUDA_TRY(cudaEventRecord(e->ev[3], st));
    unsigned long long h_q[16];
    CUDA_TRY(cudaMemcpyAsync(h_q, e->d_queue, sizeof h_q, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]));
    b->stats.kernel_ms = ms;
    b->pool_used = h_q[1];
    b->stats.dp_word_steps = h_q[2];
    b->stats.computed_cells = h_q[3];
    b->stats.passes = h_q[4];
    b->stats.fill_blocks = h_q[5];
    b->stats.dt_blocks = h_q[6];
    for (int t = 0; t < 8; t++) b->stats.phase_cycles[t] = h_q[7 + t];
    b->ran = true;
    for (uint64_t p = 0; p < b->n_pairs; p++) {
        if (b->h_status[p] == ST_BAD_INPUT) return set_err(APA_ERR_BAD_INPUT, "input byte outside ACGT in pair " + std::to_string(p));
        if (b->h_status[p] != ST_DONE)
            return set_err(APA_ERR_INTERNAL, "device assertion in pair " + std::to_string(p) + " (status " + std::to_string(b->h_status[p]) + ")");
    }
    return APA_OK;
}

extern "C" int apa_batch_run(apa_engine* e, apa_batch* b, int preset, int trace) { return batch_run(e, b, preset, trace, false); }

extern "C" int apa_batch_download(apa_engine* e, apa_batch* b, int64_t* costs, char** cigar_pool, int64_t* cigar_off, int64_t* cigar_len) {
    if (!e || !b || !b->ran) return set_err(APA_ERR_BAD_INPUT, "batch has not been run");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    if (cigar_pool) *cigar_pool = nullptr;
    if (b->n_pairs == 0) return APA_OK;
    CUDA_TRY(cudaEventRecord(e->ev[4], st));
    std::vector<int32_t> c32(b->n_pairs);
    CUDA_TRY(cudaMemcpyAsync(c32.data(), b->d_cost, b->n_pairs * 4, cudaMemcpyDeviceToHost, st));
    uint64_t bytes = b->n_pairs * 4;
    char* pool = nullptr;
    if (b->trace && cigar_pool && cigar_off && cigar_len) {
        pool = (char*)malloc(std::max<uint64_t>(b->pool_used, 1));
        if (!pool) return set_err(APA_ERR_TOO_LARGE, "host allocation of the CIGAR pool failed");
        CUDA_TRY(cudaMemcpyAsync(pool, b->d_pool, b->pool_used, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cigar_off, b->d_cig_off, b->n_pairs * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cigar_len, b->d_cig_len, b->n_pairs * 8, cudaMemcpyDeviceToHost, st));
        bytes += b->pool_used + b->n_pairs * 16;
    }
    CUDA_TRY(cudaEventRecord(e->ev[5], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (uint64_t p = 0; p < b->n_pairs; p++) costs[p] = c32[p];
    if (pool) *cigar_pool = pool;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]));
    b->stats.d2h_ms = ms;
    b->stats.d2h_bytes = bytes;
    return APA_OK;
}

extern "C" int apa_batch_get_stats(apa_batch* b, apa_batch_stats* out) {
    if (!b || !out) return APA_ERR_BAD_INPUT;
    *out = b->stats;
    return APA_OK;
}

extern "C" int apa_align_batch(apa_engine* e, int preset, int trace, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off,
                               const uint8_t* b_all, const int64_t* b_off, int64_t* costs, char** cigar_pool, int64_t* cigar_off,
                               int64_t* cigar_len, apa_batch_stats* stats) {
    // Host buffers in, host buffers out: the bases stream to HBM in chunks while the persistent kernel is already
    // aligning the pairs that have arrived (H2D overlaps compute); results come back in one D2H at the end.
    apa_batch* b = nullptr;
    int rc = batch_prepare(e, n_pairs, a_all, a_off, b_all, b_off, true, &b);
    if (rc == APA_OK) rc = batch_run(e, b, preset, trace, true);
    if (rc == APA_OK) rc = apa_batch_download(e, b, costs, cigar_pool, cigar_off, cigar_len);
    if (rc == APA_OK && stats) *stats = b->stats;
    apa_batch_free(e, b);
    return rc;
}

// ------------------------------------------------------------------------------------------------ band log (tests)
extern "C" int64_t apa_debug_band_log(apa_engine* e, int preset, int trace, const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m,
                                      int32_t* out, uint64_t cap) {
    int64_t a_off[2] = {0, (int64_t)n}, b_off[2] = {0, (int64_t)m};
    apa_batch* bt = nullptr;
    int rc = apa_batch_upload(e, 1, a, a_off, b, b_off, &bt);
    if (rc != APA_OK) return rc;
    const uint32_t rec_cap = 7u * 64u * (uint32_t)(n / BLOCK_W + 2);
    std::vector<int32_t> rec(rec_cap);
    uint32_t nrec = 0;
    cudaError_t ce = cudaMalloc(&bt->d_dbg, rec_cap * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&bt->d_dbg_n, 4);
    if (ce == cudaSuccess) ce = cudaMemset(bt->d_dbg_n, 0, 4);
    bt->dbg_cap = rec_cap;
    if (ce == cudaSuccess) rc = apa_batch_run(e, bt, preset, trace);
    if (ce == cudaSuccess && rc == APA_OK) ce = cudaMemcpy(&nrec, bt->d_dbg_n, 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && rc == APA_OK) ce = cudaMemcpy(rec.data(), bt->d_dbg, std::min(nrec, rec_cap) * 4, cudaMemcpyDeviceToHost);
    cudaFree(bt->d_dbg);
    cudaFree(bt->d_dbg_n);
    apa_batch_free(e, bt);
    if (ce != cudaSuccess) return set_err(APA_ERR_CUDA, cudaGetErrorString(ce));
    if (rc != APA_OK) return rc;
    if (nrec > rec_cap) return set_err(APA_ERR_TOO_LARGE, "band log truncated");
    // regroup into the oracle's layout: npass, then per pass: f_max, nblocks, nblocks x (j_s, j_e, fixed_s, fixed_e)
    std::vector<int32_t> o;
    o.push_back(0);
    int cur_pass = -1;
    size_t cnt_pos = 0;
    for (uint32_t r = 0; r + 7 <= nrec; r += 7) {
        if (rec[r] != cur_pass) {
            cur_pass = rec[r];
            o[0]++;
            o.push_back(rec[r + 1]);
            cnt_pos = o.size();
            o.push_back(0);
        }
        o[cnt_pos]++;
        for (int k = 3; k < 7; k++) o.push_back(rec[r + k]);
    }
    for (size_t t = 0; t < o.size() && t < cap; t++) out[t] = o[t];
    return (int64_t)o.size();
}

// ------------------------------------------------------------------------------------------------ block KAT entry
extern "C" int apa_block_compute(apa_engine* e, const uint8_t* a, uint64_t na, const uint8_t* b, uint64_t mb, uint8_t* h, uint64_t* v,
                                 int64_t* bottom_sum) {
    if (!e) return set_err(APA_ERR_NO_DEVICE, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    for (uint64_t i = 0; i < na; i++)
        if (h[i] != 1) return set_err(APA_ERR_BAD_INPUT, "apa_block_compute: top deltas must all be +1 (HMode::None)");
    const uint64_t nwords = (mb + 63) / 64, nhw = nwords * 2;
    if (na == 0 || nhw == 0) {
        *bottom_sum = (int64_t)na;
        return APA_OK;
    }
    // host-side profile of b in the device layout (same as dev_pack_planes)
    std::vector<uint2> bp(nhw);
    for (uint64_t hw = 0; hw < nhw; hw++) {
        uint32_t b0 = 0, b1 = 0;
        for (int t = 0; t < 32; t++) {
            uint64_t j = hw * 32 + t;
            if (j < mb) {
                uint32_t c = b[j];
                if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) return set_err(APA_ERR_BAD_INPUT, "byte outside ACGT");
                uint32_t x = (c >> 1) & 3u, r = x ^ (x >> 1);
                b0 |= ((r & 1u) ^ 1u) << t;
                b1 |= ((r >> 1) ^ 1u) << t;
            }
        }
        bp[hw] = make_uint2(b0, b1);
    }
    for (uint64_t i = 0; i < na; i++)
        if (!(a[i] == 'A' || a[i] == 'C' || a[i] == 'G' || a[i] == 'T')) return set_err(APA_ERR_BAD_INPUT, "byte outside ACGT");
    std::vector<uint2> vv(nhw);
    for (uint64_t w = 0; w < nwords; w++) {
        vv[2 * w] = make_uint2((uint32_t)v[2 * w], (uint32_t)v[2 * w + 1]);
        vv[2 * w + 1] = make_uint2((uint32_t)(v[2 * w] >> 32), (uint32_t)(v[2 * w + 1] >> 32));
    }
    uint8_t* d_a;
    uint2 *d_bp, *d_v, *d_fill;
    int32_t* d_cum;
    CUDA_TRY(cudaMalloc(&d_a, na));
    CUDA_TRY(cudaMalloc(&d_bp, nhw * 8));
    CUDA_TRY(cudaMalloc(&d_v, nhw * 8));
    CUDA_TRY(cudaMalloc(&d_cum, (nhw + 1) * 4));
    CUDA_TRY(cudaMalloc(&d_fill, na * nhw * 8));
    CUDA_TRY(cudaMemcpy(d_a, a, na, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_bp, bp.data(), nhw * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_v, vv.data(), nhw * 8, cudaMemcpyHostToDevice));
    apa_block_kernel<<<1, 32, 0, e->stream>>>(d_a, (int)na, d_bp, (int)nhw, d_v, d_cum, d_fill);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    std::vector<uint2> fill(na * nhw);
    std::vector<uint2> vin = vv;
    CUDA_TRY(cudaMemcpy(vv.data(), d_v, nhw * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(fill.data(), d_fill, na * nhw * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_a);
    cudaFree(d_bp);
    cudaFree(d_v);
    cudaFree(d_cum);
    cudaFree(d_fill);
    for (uint64_t w = 0; w < nwords; w++) {
        v[2 * w] = (uint64_t)vv[2 * w].x | ((uint64_t)vv[2 * w + 1].x << 32);
        v[2 * w + 1] = (uint64_t)vv[2 * w].y | ((uint64_t)vv[2 * w + 1].y << 32);
    }
    // bottom deltas: D[i][bot] - D[i-1][bot] = 1 + sum_col(i) - sum_col(i-1), with top-row deltas +1.
    auto colsum = [&](const uint2* col) {
        int64_t s = 0;
        for (uint64_t hw = 0; hw < nhw; hw++) s += __builtin_popcount(col[hw].x) - __builtin_popcount(col[hw].y);
        return s;
    };
    int64_t prev = colsum(vin.data()), total = 0;
    for (uint64_t i = 0; i < na; i++) {
        int64_t cur = colsum(&fill[i * nhw]);
        int64_t d = 1 + cur - prev;
        h[i] = d == 1 ? 1 : (d == -1 ? 2 : 0);
        total += d;
        prev = cur;
    }
    *bottom_sum = total;
    return APA_OK;
}

// ------------------------------------------------------------------------------------------------ drop-in symbols
static std::mutex g_default_mu;
static apa_engine* g_default_engine = nullptr;

static apa_engine* default_engine() {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_engine) {
        int rc = apa_engine_create(0, &g_default_engine);
        if (rc != APA_OK) {
            // Same contract as the reference, whose failures are Rust panics that abort the process
            // (astarpa-c/src/lib.rs has no error path). There is no CPU fallback.
            fprintf(stderr, "libastarpa_c (B200): %s\n", apa_last_error());
            abort();
        }
    }
    return g_default_engine;
}

static uint64_t align_one(int preset, const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                          uintptr_t* cigar_len) {
    apa_engine* e = default_engine();
    std::lock_guard<std::mutex> lk(g_default_mu);  // one stream per engine: serialise concurrent callers
    int64_t a_off[2] = {0, (int64_t)a_len}, b_off[2] = {0, (int64_t)b_len};
    int64_t cost = -1, coff = 0, clen = 0;
    char* pool = nullptr;
    int rc = apa_align_batch(e, preset, 1, 1, a, a_off, b, b_off, &cost, &pool, &coff, &clen, nullptr);
    if (rc != APA_OK) {
        fprintf(stderr, "libastarpa_c (B200): %s\n", apa_last_error());
        abort();
    }
    uint8_t* out = (uint8_t*)malloc((size_t)clen + 1);
    memcpy(out, pool + coff, (size_t)clen);
    out[clen] = 0;
    free(pool);
    *cigar_ptr = out;
    *cigar_len = (uintptr_t)clen;
    return (uint64_t)cost;
}

extern "C" uint64_t astarpa2_simple(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                                    uintptr_t* cigar_len) {
    return align_one(APA_PRESET_SIMPLE, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
extern "C" uint64_t astarpa2_full(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                                  uintptr_t* cigar_len) {
    return align_one(APA_PRESET_FULL, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
// A*PA v1 entry points (astarpa-c/src/lib.rs:54-95): served by the A*PA2 engine — same optimal cost, a valid CIGAR.
extern "C" uint64_t astarpa(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uint8_t** cigar_ptr,
                            uintptr_t* cigar_len) {
    return align_one(APA_PRESET_FULL, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
extern "C" uint64_t astarpa_gcsh(const uint8_t* a, uintptr_t a_len, const uint8_t* b, uintptr_t b_len, uintptr_t, uintptr_t, bool,
                                 uint8_t** cigar_ptr, uintptr_t* cigar_len) {
    return align_one(APA_PRESET_FULL, a, a_len, b, b_len, cigar_ptr, cigar_len);
}
extern "C" void astarpa_free_cigar(uint8_t* cigar) { free(cigar); }
