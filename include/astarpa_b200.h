/* B200-native batch extension of the A*PA2 C-ABI.
 *
 * The reference's C-ABI (astarpa-c/astarpa.h:15-65, one pair per call) cannot feed a GPU, so the single-pair
 * symbols in include/astarpa.h are served by a batch engine whose entry points are declared here. Plain C:
 * pointers and sizes only, no torch / CUDA types in any signature.
 */
#ifndef ASTARPA_B200_H
#define ASTARPA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- engine */
#define APA_PRESET_SIMPLE 0 /* AstarPa2Params::simple(), astarpa2/src/params.rs:70-96  */
#define APA_PRESET_FULL 1   /* AstarPa2Params::full(),   astarpa2/src/params.rs:98-128 */

/* Error codes (negative return values). The reference has no error convention (a Rust panic aborts the
 * process, astarpa-c/src/lib.rs:17-23); we return codes instead and never a wrong cost. */
#define APA_OK 0
#define APA_ERR_NO_DEVICE -1    /* CUDA runtime/driver/device unavailable: there is NO CPU fallback */
#define APA_ERR_CUDA -2         /* a CUDA call failed; see apa_last_error() */
#define APA_ERR_BAD_INPUT -3    /* byte outside ACGT (BitProfile::build panics, pa-bitpacking/src/profile.rs:113) */
#define APA_ERR_INTERNAL -4     /* device-side assertion (a reference panic path) */
#define APA_ERR_TOO_LARGE -5    /* sequence length >= 2^31 (I = i32) or memory budget exceeded */

typedef struct apa_engine apa_engine; /* one per GPU; owns stream, scratch arenas, result pools */
typedef struct apa_batch apa_batch;   /* a batch of pairs resident in HBM */

typedef struct apa_batch_stats {
    double h2d_ms, kernel_ms, d2h_ms; /* CUDA-event timings of the last run of this batch */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t computed_cells; /* 64 * lanes * cols evaluated by the block-DP kernel (reference: BlockStats.computed_lanes,
                                astarpa2/src/blocks.rs:705-712, times i_range.len()) */
    uint64_t dp_word_steps;  /* 32-row word x column steps issued by the block-DP kernel */
    uint64_t passes;         /* sum over pairs of f_max tries (AstarPa2Stats.f_max_tries, domain.rs:36) */
    uint64_t kernel_launches;
    uint64_t retries;        /* pairs re-run with a larger scratch arena */
    uint64_t fill_blocks, dt_blocks; /* traceback: blocks re-filled / solved by DT-trace (TraceStats, trace.rs:3-14) */
    /* SM-clock cycles summed over warps: [0] heuristic build [1] block DP [2] passes total [3] traceback total
     * [4] DT-trace [5] CIGAR text [6] h() queries [7] match pruning + contour rebuilds */
    uint64_t phase_cycles[8];
    /* CUDA-event durations of the three phase kernels of the last run: [0] apa_phase_build_kernel (GCSH build)
     * [1] apa_phase_pass_kernel (band doubling + block DP) [2] apa_phase_trace_kernel (traceback + CIGAR text).
     * All zero when the fused single-kernel path ran (arenas of the whole batch did not fit in HBM). */
    double phase_ms[3];
    /* GCSH queries of the pass kernel: h() calls and 32-layer probe rounds of the contour search (rounds / calls ~ 1 when
     * the per-stream extrapolation of the search start is good). Zero for astarpa2_simple and on the fused path. */
    uint64_t score_calls, score_probes;
    /* Which path the last run took (tests assert the shape they mean to cover):
     * pass_warps_per_pair: 1 = apa_phase_pass_kernel (one warp per pair), 4 / 8 = apa_phase_pass_coop_kernel, 0 = fused or general kernel;
     * upload_mode: 0 resident batch (apa_batch_upload), 1 host-packed planes, plain copy, 2 host-packed planes streamed under
     * the running kernel, 3 raw bases by DMA + device-side K0, plain copy, 4 raw bases streamed under the running kernel,
     * 5 streamed by both producers (raw chunks by DMA, host-packed chunks: upload_chunks_raw of upload_chunks went raw),
     * 6 / 7 caller-packed planes (apa_align_batch_packed), plain / streamed; upload_chunks: H2D chunks of the bases; waves: arena waves of the phase-split path (1 = all at once). */
    uint32_t pass_warps_per_pair, upload_mode, upload_chunks, waves;
    uint64_t dp_issue_steps; /* 32-row lane-steps ISSUED by the block DP: 32 lanes x anti-diagonals swept by every chunk (ramps, idle lanes
                                and the feeder lane included); dp_word_steps / dp_issue_steps = lane utilisation */
    uint32_t upload_chunks_raw;
    uint32_t overlapped; /* 1: the three phase kernels of the last run were launched together and overlapped at their tails
                            (phase_ms then holds overlapping intervals: build = its own run, pass / trace = 0) */
} apa_batch_stats;

const char* apa_last_error(void);
int apa_device_count(void);

int apa_engine_create(int device, apa_engine** out);
/* The process-wide engine of a device - the one apa_align_batch_multi and the single-pair symbols run on - for callers that keep
 * resident batches next to those calls (one scratch arena instead of two). Owned by the library: do not destroy it, and do
 * not use it from two threads at once. */
int apa_shared_engine(int device, apa_engine** out);
/* One process (or thread) per GPU: bind the calling thread - and with it the packing threads the engine spawns from it and the
 * page-locked buffers it touches first - to the CPUs of the GPU's NUMA node (sysfs local_cpulist of the device, intersected
 * with the caller's current affinity mask). Returns the number of CPUs bound to, 0 if nothing was changed, < 0 on error. */
int apa_bind_host_thread_to_device(int device);
void apa_engine_destroy(apa_engine* e);

/* Host -> HBM: copies the concatenated sequences (a_off/b_off have n_pairs+1 entries) and validates ACGT. */
int apa_batch_upload(apa_engine* e, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off, const uint8_t* b_all,
                     const int64_t* b_off, apa_batch** out);
/* Run the hot path on a resident batch; results stay in HBM. trace != 0 also produces CIGARs
 * (Aligner::align with trace = true, astarpa2/src/lib.rs:210-215; trace = false is AstarPa2::cost, lib.rs:177-179). */
int apa_batch_run(apa_engine* e, apa_batch* b, int preset, int trace);
/* HBM -> host: costs[n_pairs]; if the three cigar arguments are non-NULL also the CIGAR text pool: *cigar_pool is
 * callee-allocated page-locked host memory (release it with apa_free, which recycles it); pair p's NUL-terminated text starts at (*cigar_pool)[cigar_off[p]] and has
 * strlen cigar_len[p] (pairs finish in any order, so offsets are not monotone). */
int apa_batch_download(apa_engine* e, apa_batch* b, int64_t* costs, char** cigar_pool, int64_t* cigar_off, int64_t* cigar_len);
int apa_batch_get_stats(apa_batch* b, apa_batch_stats* out);
/* Per-pair counters of the last run (the stats surface of AstarPa2StatsAligner::align_with_stats, astarpa2/src/lib.rs:200-208):
 * AstarPa2Stats.f_max_tries (domain.rs:36), h0 = h(0,0) (lib.rs:124), pa_heuristic's num_matches and h_calls (GCSH only, 0
 * otherwise), BlockStats-style computed cells (64 * lanes * cols actually evaluated here; with incremental doubling the kept
 * rows and reused blocks of later passes are not counted, as in the reference), TraceStats.dt_trace_success and fill_tries (trace.rs:3-14). Timers are
 * per batch (apa_batch_stats), not per pair. */
typedef struct apa_pair_stats {
    int64_t f_max_tries, h0, num_matches, h_calls, computed_cells, dt_trace_success, fill_tries, reserved;
} apa_pair_stats;
int apa_batch_download_pair_stats(apa_engine* e, apa_batch* b, apa_pair_stats* out /* n_pairs entries */);
void apa_batch_free(apa_engine* e, apa_batch* b);
void apa_free(void* p);
/* Page-locked host buffers for the end-to-end path (H2D/D2H at full PCIe rate). */
void* apa_pinned_alloc(uint64_t bytes);
void apa_pinned_free(void* p);

/* Host buffers in, host buffers out: upload + run + download in one call (the end-to-end path). Batches of more than one
 * upload chunk (~32 MB of bases) stream: the kernels start on the first chunk while later chunks are still in flight.
 * Page-locked inputs (apa_pinned_alloc, cudaHostAlloc, cudaHostRegister) are fed by two producers taking the next chunk in turn:
 * the copy engines send raw bytes, packed to 2-bit planes on the device (K0 = BitProfile::build, pa-bitpacking/src/profile.rs:112-133,
 * inside the kernel that opens each pair), host threads pack chunks themselves (4x fewer PCIe bytes); the split follows
 * the PCIe rate and the host cores the engine finds. Pageable inputs are packed by host threads into a pinned staging buffer. */
int apa_align_batch(apa_engine* e, int preset, int trace, uint64_t n_pairs, const uint8_t* a_all, const int64_t* a_off,
                    const uint8_t* b_all, const int64_t* b_off, int64_t* costs, char** cigar_pool, int64_t* cigar_off,
                    int64_t* cigar_len, apa_batch_stats* stats);
/* The same over several GPUs of one box: pairs are independent, so the batch is cut into n_devices contiguous
 * shards balanced by bases, shard d runs on devices[d] (process-wide engine of that device, one host thread per device, no
 * collective on the data path), and the results come back in input order with ONE page-locked CIGAR pool (release with
 * apa_free). stats: NULL or n_devices entries (per-device timings of its shard). */
int apa_align_batch_multi(const int* devices, int n_devices, int preset, int trace, uint64_t n_pairs, const uint8_t* a_all,
                          const int64_t* a_off, const uint8_t* b_all, const int64_t* b_off, int64_t* costs, char** cigar_pool,
                          int64_t* cigar_off, int64_t* cigar_len, apa_batch_stats* stats);
/* Packed (2-bit) input: for pipelines that keep their reads packed, and for boxes where ASCII bases through host memory are
 * the limit (8 GPUs x 2 GB of bases per step). A packed sequence set is the engine's own plane layout: sequence p occupies the
 * half-words [off[p], off[p + 1]) (apa_packed_layout: ceil(len / 64) * 2 + 2 half-words rounded up to a multiple of 16); a
 * half-word is two u32, (plane0, plane1) = the NEGATED rank bits (A0 C1 G2 T3) of 32 consecutive bases, zero past the end of
 * the sequence (BitProfile::build, pa-bitpacking/src/profile.rs:124-131). apa_pack_sequences fills such an array on host threads
 * (validating ACGT); apa_align_batch_packed is apa_align_batch on such arrays: the planes go to HBM by DMA alone. */
int apa_packed_layout(uint64_t n, const int64_t* len, int64_t* off_out /* n + 1 entries, in half-words */);
int apa_pack_sequences(uint64_t n, const uint8_t* seq_all, const int64_t* seq_off /* n + 1 */, uint32_t* planes_out,
                       const int64_t* off /* from apa_packed_layout */, int n_threads);
int apa_align_batch_packed(apa_engine* e, int preset, int trace, uint64_t n_pairs, const uint32_t* a_planes, const int64_t* a_len,
                           const uint32_t* b_planes, const int64_t* b_len, int64_t* costs, char** cigar_pool, int64_t* cigar_off,
                           int64_t* cigar_len, apa_batch_stats* stats);

/* K0 on its own (tests): BitProfile::build of one sequence on the device, layout as in the engine - per 32 bases one
 * (plane0, plane1) pair of u32 with the NEGATED rank bits of A0 C1 G2 T3, ceil(len / 64) * 2 + 2 half-words rounded up to
 * a multiple of 16, zero past the end. out receives 2 u32 per half-word; *n_halfwords the count. Returns APA_ERR_BAD_INPUT for
 * a byte outside ACGT (the reference panics, profile.rs:113). */
int apa_pack_planes_device(apa_engine* e, const uint8_t* seq, uint64_t len, uint32_t* out, uint64_t out_cap_halfwords,
                           uint64_t* n_halfwords);

/* ---------------------------------------------------------------- general parameters
 * AstarPa2Params (astarpa2/src/params.rs:8-42) beyond the two presets: the other Domains and DoublingTypes, block widths,
 * heuristics and BlockParams knobs of the reference's own test matrix (astarpa2/src/tests.rs:19-119). Served by a second,
 * general kernel (same device code, parameters read at run time); results are bit-exact against the oracle like the presets.
 * Not supported (APA_ERR_BAD_INPUT): the SH / CSH heuristics, inexact matches (r = 2), sparse = false, LocalDoubling
 * (ignored as broken in the reference, tests.rs:121-130), viz. */
#define APA_DOMAIN_FULL 0      /* Domain::Full: the whole rectangle (params.rs:232) */
#define APA_DOMAIN_GAP_START 1 /* states with gap(s, u) <= f (params.rs:234) */
#define APA_DOMAIN_GAP_GAP 2   /* gap(s, u) + gap(u, t) <= f, Edlib-like (params.rs:236) */
#define APA_DOMAIN_ASTAR 3     /* g(u) + h(u) <= f (params.rs:240) */
#define APA_HEURISTIC_NONE 0   /* NoCost: Dijkstra */
#define APA_HEURISTIC_GAP 1    /* GapCost (pa-heuristic/src/heuristic/distances.rs:130-169) */
#define APA_HEURISTIC_GCSH 2   /* GCSH with exact matches of length k, Prune::Start (pa-heuristic/src/heuristic/csh.rs) */
#define APA_DOUBLING_NONE 0    /* one unbounded pass; requires APA_DOMAIN_FULL (lib.rs:126-130) */
#define APA_DOUBLING_BAND 1    /* DoublingType::BandDoubling{start, factor} (band.rs:100-141) */
#define APA_DOUBLING_LINEAR 2  /* DoublingType::LinearSearch{start, delta} (band.rs:142-190) */
#define APA_START_ZERO 0       /* DoublingStart (band.rs:4-23) */
#define APA_START_GAP 1
#define APA_START_H0 2
typedef struct apa_params {
    int32_t domain;
    int32_t heuristic;    /* read when domain == APA_DOMAIN_ASTAR */
    int32_t k;            /* GCSH seed length, 4..16 (HeuristicParams.k) */
    int32_t r;            /* must be 1 (exact matches) */
    int32_t p;            /* GCSH local-pruning look-ahead in seeds, 0 (off) ..15 (HeuristicParams.p) */
    int32_t doubling;
    int32_t doubling_start;
    float factor;         /* BandDoubling */
    int32_t delta;        /* LinearSearch */
    int32_t block_width;  /* 1..256 */
    int32_t sparse;       /* BlockParams.sparse: must be 1 */
    int32_t incremental_doubling; /* blocks.rs:342-469: later passes keep the rows the previous pass had fixed and reuse unchanged blocks */
    int32_t dt_trace;
    int32_t max_g;        /* 1..40 */
    int32_t fr_drop;      /* 0 disables the x-drop */
    int32_t sparse_h;
    int32_t prune;
} apa_params;
/* Fills *out with AstarPa2Params::simple() / ::full() (params.rs:70-128). */
int apa_params_preset(int preset, apa_params* out);
/* AstarPa2Params from the serde JSON the reference and pa-bench exchange (params.rs:7-42; missing fields take serde's defaults,
 * unknown fields are an error like #[serde(deny_unknown_fields)]). Values this engine does not serve are refused with
 * APA_ERR_BAD_INPUT and a message naming the field in err (err_cap bytes, may be NULL). */
int apa_params_from_json(const char* json, apa_params* out, char* err, uint64_t err_cap);
/* apa_batch_run / apa_align_batch / apa_debug_band_log with explicit parameters (always the general kernel). */
int apa_batch_run_params(apa_engine* e, apa_batch* b, const apa_params* params, int trace);
int apa_align_batch_params(apa_engine* e, const apa_params* params, int trace, uint64_t n_pairs, const uint8_t* a_all,
                           const int64_t* a_off, const uint8_t* b_all, const int64_t* b_off, int64_t* costs, char** cigar_pool,
                           int64_t* cigar_off, int64_t* cigar_len, apa_batch_stats* stats);
int64_t apa_debug_band_log_params(apa_engine* e, const apa_params* params, int trace, const uint8_t* a, uint64_t n,
                                  const uint8_t* b, uint64_t m, int32_t* out, uint64_t cap);

/* Debug/test introspection: per-pass band log of one pair in the oracle's layout
 * (passes, then per pass: f_max, nblocks, nblocks x (j_s, j_e, fixed_s, fixed_e)). Returns int32 count or <0. */
int64_t apa_debug_band_log(apa_engine* e, int preset, int trace, const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m,
                           int32_t* out, uint64_t cap);

/* pa_bitpacking::search (pa-bitpacking/src/search.rs:46-118): semi-global search of a short pattern (may contain the
 * wildcards N, *, Y = C|T, R = A|G) in a long text (acgtACGT). A match may start anywhere in the text and anywhere in the
 * pattern; unmatched pattern rows cost `unmatched_cost` each (0..1, realised as one +1 row every 1/unmatched_cost rows).
 * out receives pattern_len + text_len + 1 costs: along the bottom row, then up the right column (search.rs:36-45). The DP
 * runs on the GPU (the block-DP step with zero top deltas and the pattern's match masks as equality words), one warp per text
 * segment: a cell of row j costs at most j, so a warp that starts 2 * rows columns ahead of its segment from an upper bound of
 * the column is exact inside the segment. */
int apa_search(apa_engine* e, const uint8_t* pattern, uint64_t pattern_len, const uint8_t* text, uint64_t text_len,
               float unmatched_cost, int32_t* out);

/* SearchResult::trace(idx) of pa_bitpacking::search (pa-bitpacking/src/search.rs:135-230): the alignment of the pattern that ends at
 * out[idx] (idx <= text_len: bottom row, column idx; beyond: up the right column). The text window that can hold it is re-filled
 * and walked back on the GPU. cigar_out receives the NUL-terminated CIGAR text ('=' match, 'X' substitution, 'D' a text base
 * skipped, 'I' a pattern base skipped; count omitted when 1), pos_out[5] = {start.i, start.j, end.i, end.j, cost} with i the
 * text column and j the pattern row; start is where the walk stopped (the window's left edge or the top row). */
int apa_search_trace(apa_engine* e, const uint8_t* pattern, uint64_t pattern_len, const uint8_t* text, uint64_t text_len,
                     float unmatched_cost, uint64_t idx, char* cigar_out, uint64_t cigar_cap, int32_t* pos_out);

/* INT32 issue-rate probe (bench.py's roofline denominator for the block DP, SURVEY 8d "must be measured"): lane-operations
 * per second with every SM full of warps running dependent chains of out[0] LOP3, out[1] SHF, out[2] IADD3 (three-input add),
 * out[3] IMAD; out[4] = the ALU-pipe instructions (8 LOP3 + 2 SHF per Myers step) of the block-DP instruction mix with its
 * 4 IMAD per step running next to them on the FMA pipe, out[6] = those IMADs per second; out[5] = the device clock attribute in Hz. */
int apa_int32_peak(apa_engine* e, double* out /* 7 doubles */);

/* Block-DP kernel on its own (pa_bitpacking::simd::compute semantics, pa-bitpacking/src/simd.rs:98-226, HMode::Update of
 * astarpa2/src/blocks.rs:665-748): rectangle a[na] x b[mb]; h one byte per column (0, 1 = +1, 2 = -1): the deltas along the top
 * edge on entry, along the bottom edge on return; v interleaved (p,m) u64 pairs per 64-row word, in/out. *bottom_sum = sum of
 * the returned bottom deltas. */
int apa_block_compute(apa_engine* e, const uint8_t* a, uint64_t na, const uint8_t* b, uint64_t mb, uint8_t* h, uint64_t* v,
                      int64_t* bottom_sum);

#ifdef __cplusplus
}
#endif
#endif
