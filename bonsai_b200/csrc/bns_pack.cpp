// bns_pack.cpp -- ASCII bases -> 2-bit units on the host (format and purpose: bns_pack.h). AVX-512BW, AVX2 and scalar
// bodies behind a run-time dispatch; the three produce identical bytes (tests/host/pack_check.cpp).
#include "bns_pack.h"

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace bns {
namespace {

// code of a byte: the arithmetic of the device's pack4 (bns_device.cuh), so that both paths agree on EVERY byte
inline unsigned code_of(unsigned ch) { unsigned t = (ch >> 1) & 3u; return t ^ (t >> 1); }
inline bool valid_base(unsigned ch) { const unsigned up = ch & 0xdfu; return up == 'A' || up == 'C' || up == 'G' || up == 'T'; }
inline unsigned rev8(unsigned b) {                                    // base i of a unit: bit i -> bit 7 - i
    b = ((b & 0xf0u) >> 4) | ((b & 0x0fu) << 4);
    b = ((b & 0xccu) >> 2) | ((b & 0x33u) << 2);
    return ((b & 0xaau) >> 1) | ((b & 0x55u) << 1);
}

// units [u0, u1) of the stream one byte at a time; a unit past the end of the stream is padded with code 0 (never read as a k-mer)
void pack_units_scalar(const unsigned char *s, size_t n, size_t u0, size_t u1, uint16_t *units, uint32_t *susp, std::vector<uint64_t> &exc,
                       uint64_t unit0) {
    for(size_t u = u0; u < u1; ++u) {
        unsigned v = 0, bad = 0;
        for(unsigned i = 0; i < 8; ++i) {
            const size_t p = 8 * u + i;
            const unsigned ch = p < n ? s[p] : (unsigned)'A';
            v = (v << 2) | code_of(ch);
            bad = (bad << 1) | (valid_base(ch) ? 0u : 1u);
        }
        units[u] = (uint16_t)v;
        if((u & 31u) == 0) susp[u >> 5] = 0;
        if(bad) { susp[u >> 5] |= 1u << (u & 31u); exc.push_back(((unit0 + u) << 8) | bad); }
    }
}

// the units of 64 bases whose validity mask (bit i: byte i is one of ACGTacgt) is not all ones
inline uint32_t note_bad(uint64_t bad, size_t u, std::vector<uint64_t> &exc, uint64_t unit0) {
    uint32_t sw = 0;
    for(unsigned j = 0; j < 8; ++j) {
        const unsigned b = (unsigned)(bad >> (8 * j)) & 0xffu;
        if(b) { sw |= 1u << j; exc.push_back(((unit0 + u + j) << 8) | rev8(b)); }
    }
    return sw;
}

// Per 16 low-nibble values: the code (bits 1-2 of the byte decide) and the upper-case byte a valid base with that nibble
// must be ('A' 0x41, 'C' 0x43, 'T' 0x54, 'G' 0x47; 0xff never equals a byte with bit 5 cleared).
#define BNS_LUT_CODE 0, 0, 1, 1, 3, 3, 2, 2, 0, 0, 1, 1, 3, 3, 2, 2
#define BNS_LUT_EXPECT -1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1

// NT: the units leave with non-temporal stores (16-byte aligned destination): the staging buffer is read next by the copy
// engine, not by this core, and a streaming store spares the read-for-ownership of every line
template <bool NT>
__attribute__((target("avx512f,avx512bw,avx512vl")))
size_t pack_words_avx512(const unsigned char *s, size_t n_words, uint16_t *units, uint32_t *susp, std::vector<uint64_t> &exc, uint64_t unit0) {
    const __m512i lut_code = _mm512_broadcast_i32x4(_mm_setr_epi8(BNS_LUT_CODE));
    const __m512i lut_exp = _mm512_broadcast_i32x4(_mm_setr_epi8(BNS_LUT_EXPECT));
    const __m512i nib = _mm512_set1_epi8(0x0f), upper = _mm512_set1_epi8((char)0xdf);
    const __m512i w1 = _mm512_set1_epi16(0x0104);                     // bytes (4, 1): c0 * 4 + c1
    const __m512i w2 = _mm512_set1_epi32(0x00010010);                 // words (16, 1): the byte of four bases
    const __m128i swap = _mm_setr_epi8(1, 0, 3, 2, 5, 4, 7, 6, 9, 8, 11, 10, 13, 12, 15, 14);   // first four bases in the HIGH byte of the u16
    for(size_t w = 0; w < n_words; ++w) {
        uint32_t sw = 0;
        for(unsigned b = 0; b < 4; ++b) {
            const size_t u = 32 * w + 8 * b;
            const __m512i v = _mm512_loadu_si512(s + 8 * u);
            const __m512i idx = _mm512_and_si512(v, nib);
            const uint64_t ok = _mm512_cmpeq_epi8_mask(_mm512_and_si512(v, upper), _mm512_shuffle_epi8(lut_exp, idx));
            const __m512i c = _mm512_shuffle_epi8(lut_code, idx);
            const __m512i q = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w1), w2);
            const __m128i o = _mm_shuffle_epi8(_mm512_cvtepi32_epi8(q), swap);
            if(NT) _mm_stream_si128((__m128i *)(units + u), o); else _mm_storeu_si128((__m128i *)(units + u), o);
            if(ok != ~0ull) sw |= note_bad(~ok, u, exc, unit0) << (8 * b);
        }
        susp[w] = sw;
    }
    if(NT) _mm_sfence();
    return n_words;
}

__attribute__((target("avx2")))
size_t pack_words_avx2(const unsigned char *s, size_t n_words, uint16_t *units, uint32_t *susp, std::vector<uint64_t> &exc, uint64_t unit0) {
    const __m256i lut_code = _mm256_broadcastsi128_si256(_mm_setr_epi8(BNS_LUT_CODE));
    const __m256i lut_exp = _mm256_broadcastsi128_si256(_mm_setr_epi8(BNS_LUT_EXPECT));
    const __m256i nib = _mm256_set1_epi8(0x0f), upper = _mm256_set1_epi8((char)0xdf);
    const __m256i w1 = _mm256_set1_epi16(0x0104), w2 = _mm256_set1_epi32(0x00010010);
    // per 128-bit lane: the low bytes of its four dwords, pairwise swapped, into the lane's first four bytes
    const __m256i gather = _mm256_broadcastsi128_si256(_mm_setr_epi8(4, 0, 12, 8, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    for(size_t w = 0; w < n_words; ++w) {
        uint32_t sw = 0;
        for(unsigned b = 0; b < 8; ++b) {
            const size_t u = 32 * w + 4 * b;
            const __m256i v = _mm256_loadu_si256((const __m256i *)(s + 8 * u));
            const __m256i idx = _mm256_and_si256(v, nib);
            const uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(v, upper), _mm256_shuffle_epi8(lut_exp, idx)));
            const __m256i c = _mm256_shuffle_epi8(lut_code, idx);
            const __m256i q = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(c, w1), w2), gather);
            const uint32_t lo = (uint32_t)_mm256_extract_epi32(q, 0), hi = (uint32_t)_mm256_extract_epi32(q, 4);
            const uint64_t both = ((uint64_t)hi << 32) | lo;
            __builtin_memcpy(units + u, &both, 8);
            if(ok != ~0u) {
                const uint32_t bad = ~ok;
                for(unsigned j = 0; j < 4; ++j) {
                    const unsigned bb = (bad >> (8 * j)) & 0xffu;
                    if(bb) { sw |= 1u << (4 * b + j); exc.push_back(((unit0 + u + j) << 8) | rev8(bb)); }
                }
            }
        }
        susp[w] = sw;
    }
    return n_words;
}

int detect_isa() {
    __builtin_cpu_init();
    if(const char *e = std::getenv("BNS_B200_PACK_ISA")) {           // tests: force a narrower body
        if(!std::strcmp(e, "scalar")) return 0;
        if(!std::strcmp(e, "avx2")) return __builtin_cpu_supports("avx2") ? 1 : 0;
    }
    if(__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) return 2;
    if(__builtin_cpu_supports("avx2")) return 1;
    return 0;
}
int isa() { static const int v = detect_isa(); return v; }

}  // namespace

const char *pack_isa() { return isa() == 2 ? "avx512" : isa() == 1 ? "avx2" : "scalar"; }

void pack_range(const char *bases, size_t n, uint16_t *units, uint32_t *susp, std::vector<uint64_t> &exc, uint64_t unit0) {
    const unsigned char *s = (const unsigned char *)bases;
    const size_t n_words = n / 256;                                   // whole suspicious-bit words (32 units) go the wide way
    size_t done = 0;
    static const bool nt_ok = !std::getenv("BNS_B200_PACK_NO_NT");
    if(isa() == 2) done = (nt_ok && ((uintptr_t)units & 15u) == 0) ? pack_words_avx512<true>(s, n_words, units, susp, exc, unit0)
                                                                  : pack_words_avx512<false>(s, n_words, units, susp, exc, unit0);
    else if(isa() == 1) done = pack_words_avx2(s, n_words, units, susp, exc, unit0);
    pack_units_scalar(s, n, done * 32, (n + 7) / 8, units, susp, exc, unit0);
}

// ---- worker threads ----------------------------------------------------------------------------------------------------
struct PackPool::Impl {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, cv_done;
    std::function<void(unsigned)> fn;
    unsigned n_tasks = 0, active = 0;
    std::atomic<unsigned> next{0};
    std::atomic<uint64_t> gen{0};
    std::atomic<bool> finished{true};
    bool stop = false;

    void work() {
        uint64_t seen = 0;
        for(;;) {
            // a short spin first: jobs follow one another within microseconds while a call is running
            for(int i = 0; i < 4000 && gen.load(std::memory_order_acquire) == seen; ++i) _mm_pause();
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return stop || gen.load(std::memory_order_acquire) != seen; });
                if(stop) return;
                seen = gen.load(std::memory_order_acquire);
            }
            for(;;) {
                const unsigned t = next.fetch_add(1, std::memory_order_relaxed);
                if(t >= n_tasks) break;
                fn(t);
            }
            std::lock_guard<std::mutex> lk(m);
            if(--active == 0) { finished.store(true, std::memory_order_release); cv_done.notify_all(); }
        }
    }
};

PackPool::PackPool(unsigned n_threads) : n_(n_threads ? n_threads : 1), impl_(new Impl) {
    for(unsigned i = 0; i < n_; ++i) impl_->th.emplace_back([this] { impl_->work(); });
}
PackPool::~PackPool() {
    wait();
    { std::lock_guard<std::mutex> lk(impl_->m); impl_->stop = true; }
    impl_->cv.notify_all();
    for(auto &t : impl_->th) t.join();
    delete impl_;
}
void PackPool::start(unsigned n_tasks, std::function<void(unsigned)> fn) {
    wait();
    std::lock_guard<std::mutex> lk(impl_->m);
    impl_->fn = std::move(fn);
    impl_->n_tasks = n_tasks;
    impl_->next.store(0);
    impl_->active = n_;
    impl_->finished.store(false, std::memory_order_release);
    impl_->gen.fetch_add(1, std::memory_order_release);
    impl_->cv.notify_all();
}
bool PackPool::done() const { return impl_->finished.load(std::memory_order_acquire); }
void PackPool::wait() {
    if(done()) return;
    std::unique_lock<std::mutex> lk(impl_->m);
    impl_->cv_done.wait(lk, [&] { return impl_->finished.load(std::memory_order_acquire); });
}

}  // namespace bns
