// bns_host_util.h -- host-only helpers of the loader (plain C++, no CUDA): kept apart so that tests/host/ can exercise them
// without a device.
#pragma once
#include <algorithm>
#include <cstdint>
#include <iterator>
#include <vector>

namespace bns {

// The sorted distinct values of a stream of u32 (the DB's taxids: the value dictionary of the device table). Values below
// 2^24 -- every NCBI taxid so far -- are marked in a 2 MiB bitmap, one test-and-set per key instead of sorting all of
// them (10.5 M keys: 0.54 s -> 0.03 s); larger values are collected and sort-merged in batches.
class DistinctValues {
    static constexpr uint32_t LOW_BITS = 24;
    std::vector<uint64_t> bitmap_;
    std::vector<uint32_t> high_, pending_;
    void fold() {
        std::sort(pending_.begin(), pending_.end());
        pending_.erase(std::unique(pending_.begin(), pending_.end()), pending_.end());
        std::vector<uint32_t> merged;
        merged.reserve(high_.size() + pending_.size());
        std::set_union(high_.begin(), high_.end(), pending_.begin(), pending_.end(), std::back_inserter(merged));
        high_.swap(merged);
        pending_.clear();
    }

public:
    DistinctValues() : bitmap_((size_t)1 << (LOW_BITS - 6), 0) {}
    void add(uint32_t v) {
        if(v >> LOW_BITS) {
            if(!pending_.empty() && pending_.back() == v) return;
            pending_.push_back(v);
            if(pending_.size() >= ((size_t)1 << 22)) fold();
        } else bitmap_[v >> 6] |= (uint64_t)1 << (v & 63);
    }
    std::vector<uint32_t> sorted() {
        fold();
        std::vector<uint32_t> out;
        for(size_t w = 0; w < bitmap_.size(); ++w)
            for(uint64_t m = bitmap_[w]; m; m &= m - 1) out.push_back((uint32_t)(w << 6) + (uint32_t)__builtin_ctzll(m));
        out.insert(out.end(), high_.begin(), high_.end());
        return out;
    }
};

inline std::vector<uint32_t> distinct_values(const uint32_t *vals, uint64_t n) {
    DistinctValues d;
    for(uint64_t i = 0; i < n; ++i) d.add(vals[i]);
    return d.sorted();
}

// The same over the occupied buckets of a khash value array (flags: 2 bits per bucket, bit1 empty / bit0 deleted, 16 buckets
// per word, khash64.h:169-177; a table of fewer than 16 buckets still has one word). *n_occupied: buckets with both bits clear.
inline std::vector<uint32_t> distinct_values_khash(const uint32_t *vals, const uint32_t *flags, uint64_t n_buckets, uint64_t *n_occupied) {
    DistinctValues d;
    uint64_t n = 0;
    for(uint64_t base = 0; base < n_buckets; base += 16) {
        const uint32_t f = flags[base >> 4];
        uint32_t occ = ~(f | (f >> 1)) & 0x55555555u;                  // bit 2j set: bucket base + j is occupied
        if(n_buckets - base < 16) occ &= ((uint32_t)1 << (2 * (n_buckets - base))) - 1u;
        for(; occ; occ &= occ - 1) {
            d.add(vals[base + ((uint32_t)__builtin_ctz(occ) >> 1)]);
            ++n;
        }
    }
    if(n_occupied) *n_occupied = n;
    return d.sorted();
}

}  // namespace bns
