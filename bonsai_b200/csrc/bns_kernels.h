// bns_kernels.h -- launcher interface between the host API (bns_api.cu) and the kernels (bns_kernels.cu)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace bns {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr int WARPS_PER_CTA = 8;

struct EncParams;
struct TableView;
struct TaxView;
struct TableFmt;

size_t stream_smem_bytes(u32 ring_cap, bool classify);
enum LeanMode : int { LEAN_U = 0, LEAN_K = 1, LEAN_R = 2, LEAN_S = 3 };   // what bns_classify_u_kernel runs as (bns_classify_u.cuh)
enum LeanKey : int { LEAN_KEY_PAIR = 0, LEAN_KEY_LEX = 1, LEAN_KEY_ELEM = 2 };   // how window elements are ordered
// host-packed bases of one chunk (bns_pack.h): one suspicious bit per 16-bit unit, sorted exception words, first base of the stream
struct PackedIn { const u32 *susp; const unsigned long long *exc; u32 n_exc; u64 base0; };
struct RunsOut { u64 *runs; u64 cap; unsigned long long *total; u64 *run_pos; u32 *n_runs; };   // run-list outputs of the lean kernel
struct ClassifyPlan { int grid = 1; size_t smem = 0; bool lean = false, counts = true, loc = false, runs = false; int occupancy = 0, lean_mode = -1, gen_grid = 1; size_t gen_smem = 0;
                      bool second_pass = false; int pass2_grid = 1; u32 big_cap = 0;   // the pass over the records the first kernel left (bns_kernels.cu)
                      u32 fixed_len = 0; u64 fixed_base = 0;   // set by the caller: all records have fixed_len bases, offsets are not on the device
                      bool sv = false;                        // value dictionary of <= 32 entries: the lean kernel counts in lane registers
                      int lean_warps = 8;                     // warps per CTA the lean kernel is launched with
                      bool packed = false;                    // `bases` is the 2-bit unit stream of bns_pack.h (launch_classify needs its PackedIn)
                    };
ClassifyPlan plan_classify(const EncParams &P, const TableView &T, u32 ring_cap, int n_sm, u64 n_records, u32 mates, bool taxa, bool mate1, bool counts,
                           bool runs = false, bool packed = false);
int encode_occupancy(const EncParams &P, size_t smem);
// sequences one full wave of the lean kernel takes (every warp of every SM one batch): chunk sizes that are multiples of it leave no
// warp idle while others run a last batch
u64 lean_wave_reads(int n_sm);
// entries of the chunk's run buffer beyond one per possible hit: every warp of the lean run-list kernel may leave a stretch unused
u64 runs_slack(const ClassifyPlan &pl);

cudaError_t launch_encode(const EncParams &P, int grid, size_t smem, cudaStream_t st, const char *bases, const u64 *offsets, u64 n_seqs,
                          u64 total_bases, u64 *kmers_out, const u64 *out_offsets, u32 *counts_out, u32 ring_cap, u32 *status);
cudaError_t launch_classify(const EncParams &P, const ClassifyPlan &pl, cudaStream_t st, const char *bases, const u64 *offsets, u64 n_records,
                            u32 mates, u64 total_bases, const TableView &T, const TaxView &X,
                            u32 *taxon_out, u32 *nhit_out, u32 *nmiss_out, u32 *taxa_out, const u64 *taxa_offsets,
                            u32 *mate1_out, u32 ring_cap, unsigned long long *counters, u32 *status,
                            u32 *defer_idx, unsigned long long *defer_cnt, int *n_launched, const RunsOut *ro = nullptr, u32 *big_scratch = nullptr,
                            const PackedIn *pk = nullptr);
// u32 words of global scratch the second pass needs (0 unless the database holds more than AGG_CAP distinct values)
size_t pass2_scratch_words(const ClassifyPlan &pl);
cudaError_t launch_build(const EncParams &P, int grid, size_t smem, cudaStream_t st, const char *bases, const u64 *offsets,
                         u64 n_seqs, u64 total_bases, u64 *slots, const TableFmt &fmt, u32 vid, const TaxView &X, const u32 *values,
                         u32 n_values, unsigned long long *stats, u32 ring_cap);
cudaError_t launch_dump(cudaStream_t st, const u64 *slots, u64 n_buckets, const TableFmt &fmt, const u32 *dict, u64 *keys_out, u32 *vals_out,
                        u64 cap, unsigned long long *counter);
cudaError_t launch_insert(cudaStream_t st, u64 *slots, const TableFmt &fmt, const u64 *keys, const u32 *vals, u64 n, const u32 *values,
                          u32 n_values, unsigned long long *stats, u64 *fail_keys = nullptr, u32 *fail_vals = nullptr, u64 fail_cap = 0);
cudaError_t launch_table_stats(cudaStream_t st, const u64 *slots, u64 n_buckets, const TableFmt &fmt, unsigned long long *out);
cudaError_t launch_lookup(cudaStream_t st, const TableView &T, const u32 *dict, const u64 *keys, u64 n, u32 *vals_out,
                          uint8_t *found_out);
cudaError_t launch_sectors(cudaStream_t st, const TableView &T, const u64 *keys, u64 n, unsigned long long *total);
cudaError_t launch_resolve(int grid, cudaStream_t st, const TaxView &X, const u32 *values, u32 n_values, const u32 *taxa,
                           const uint16_t *counts, const u64 *offsets, u64 n_lists, u32 *taxon_out, u32 *status);
cudaError_t launch_rle(cudaStream_t st, const u32 *taxa, const u64 *taxa_offsets, const u32 *nhit, u64 n_records, u64 *runs,
                       unsigned long long *total, u64 *run_pos, u32 *n_runs);
cudaError_t launch_gather(int grid, cudaStream_t st, const u64 *slots, u32 b, u64 n_loads, u64 seed, unsigned long long *sink);

}  // namespace bns
