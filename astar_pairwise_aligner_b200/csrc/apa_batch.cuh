// Kernel arguments shared by the two translation units of the engine: the preset kernels (apa_engine.cu, everything a
// compile-time constant of astarpa2_simple / astarpa2_full) and the general-parameter kernel (apa_general.cu, the same
// device code compiled with run-time AstarPa2Params).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct BatchDev {
    uint64_t n_pairs;
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
    uint32_t q0;        // phase-split path: first work-order position of this wave (arena slot = q - q0)
    // Streaming upload: chunk_state[c] = 0 while upload chunk c is on its way, 1 once its packed planes are in HBM, 2 once its raw
    // bases are (the kernel then packs them: device-side K0); pair_chunk[q] = chunk of work-order position q. nullptr = all there.
    const volatile uint32_t* chunk_state;
    const uint16_t* pair_chunk;
    long long* pair_stats;  // 8 per pair: apa_pair_stats (f_max_tries, h0, num_matches, h_calls, computed_cells, dt blocks, fill tries, -)
    // Device-side K0: when non-null, the raw bases as uploaded (byte offsets a_off / b_off); the kernel that opens a pair packs
    // them into aprof / bprof first (dev_pack_planes). Null: the planes were packed before the launch.
    const uint8_t* raw_a;
    const uint8_t* raw_b;
    // Overlapped phase kernels (one warp per pair, whole batch in one wave): the build, pass and trace kernels are launched
    // together on three streams; phase_flag[q] (work-order position) is 1 once pair q is built and 2 once its passes are done,
    // and the kernel of the next phase waits for it before it opens the pair - so its CTAs fill the SM slots the previous
    // kernel's tail leaves empty. nullptr = the kernels run back to back.
    volatile uint8_t* phase_flag;
    int32_t* dbg;  // band log of the (single) pair, or nullptr
    uint32_t dbg_cap;
    uint32_t* dbg_n;
};

constexpr int WARPS_PER_CTA = 4;

// AstarPa2Params (astarpa2/src/params.rs:8-42) as the general kernel sees them. Values mirror include/astarpa_b200.h.
struct RunParams {
    int domain;        // 0 Full, 1 GapStart, 2 GapGap, 3 Astar            (params.rs:230-242)
    int heuristic;     // 0 None (NoCost), 1 Gap (GapCost), 2 GCSH           (only read when domain == Astar)
    int k, p;          // GCSH: seed length, local-pruning look-ahead (0 = off); r = 1 exact matches only
    int doubling;      // 0 None, 1 BandDoubling, 2 LinearSearch             (band.rs:26-45)
    int start;         // DoublingStart: 0 Zero, 1 Gap, 2 H0                 (band.rs:4-23)
    float factor;      // BandDoubling growth factor
    int delta;         // LinearSearch step
    int block_width;   // 1 ..= 256
    int dt_trace, max_g, fr_drop;  // BlockParams (blocks.rs:31-74); sparse = true always
    int sparse_h, prune;
    int incremental;   // BlockParams.incremental_doubling
};
cudaError_t apa_general_launch(const BatchDev& bd, const RunParams& par, unsigned grid, cudaStream_t st);
