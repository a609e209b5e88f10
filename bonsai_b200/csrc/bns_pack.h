// bns_pack.h -- host side of the packed input path: ASCII bases -> 2-bit codes on the caller's cores, so that 38 bytes per
// 150 bp read cross PCIe instead of 150 (the e2e bound of the host-buffer calls). Plain C++ (g++), no CUDA.
//
// Format (what bns_classify_u_kernel<..., PK = true> reads): the bases of a chunk, taken as ONE stream starting at the chunk's
// first base, in 16-bit units of 8 bases (first base in the top two bits; the unit is a little-endian u16 in memory); one
// "suspicious" bit per unit (bit u & 31 of word u >> 5: the unit holds a byte that is not ACGTacgt) and, for those units
// only, an exception word (unit << 8 | invalid mask, base i of the unit at bit 7 - i), sorted by unit. Codes are the
// reference's (A 0, C 1, G 2, T 3, lower case alike: alphabet.h / kmerutil.h cstr_lut); an invalid byte packs as some code
// and is masked by its exception, exactly as the device's own pack4 treats the ASCII stream (bns_device.cuh).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace bns {

// bases[0, n) -> units[0, (n + 7) / 8), susp words [0, (n + 255) / 256) (both fully written), exceptions appended to `exc`
// with unit numbers starting at unit0. `n` must be a multiple of 256 unless this is the last piece of the stream.
void pack_range(const char *bases, size_t n, uint16_t *units, uint32_t *susp, std::vector<uint64_t> &exc, uint64_t unit0);
// name of the instruction set pack_range dispatches to on this machine ("avx512", "avx2", "scalar")
const char *pack_isa();

// a fixed set of worker threads that run fn(0..n_tasks-1); one job at a time, started asynchronously
class PackPool {
  public:
    explicit PackPool(unsigned n_threads);
    ~PackPool();
    unsigned size() const { return n_; }
    void start(unsigned n_tasks, std::function<void(unsigned)> fn);   // returns at once; the job runs on the workers
    bool done() const;                                                // the job started last has finished (true when none was)
    void wait();
    struct Impl;

  private:
    unsigned n_;
    Impl *impl_;
};

}  // namespace bns
