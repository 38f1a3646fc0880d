// Library-level exports: version, last error, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include <thread>
#include <vector>

#include "common.cuh"

namespace sym {
static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- content digest of host arrays (cache keys of the Python layer: has the caller edited the reference?) ----
// Eight independent 64-bit lanes per 64-byte block (add-multiply, rotate, multiply: one multiply chain per lane keeps
// a core at memory speed), folded with the length at the end.  Change detection, not cryptography.
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t avalanche(uint64_t h) {
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    h *= 0x9FB21C651E98DF25ull;
    h ^= h >> 32;
    return h;
}
static uint64_t digest_span(const unsigned char *p, uint64_t n, uint64_t seed) {
    constexpr uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full;
    uint64_t acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = seed + (uint64_t)(i + 1) * P2;
    uint64_t left = n;
    while (left >= 64) {
        uint64_t w[8];
        memcpy(w, p, 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = rotl64(acc[i] + w[i] * P1, 27) * P2;
        p += 64;
        left -= 64;
    }
    if (left) {
        uint64_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        memcpy(w, p, left);
        for (int i = 0; i < 8; ++i) acc[i] = rotl64(acc[i] + w[i] * P1, 27) * P2;
    }
    uint64_t h = n * P1 + seed;
    for (int i = 0; i < 8; ++i) h = rotl64(h ^ avalanche(acc[i]), 23) * P1 + P2;
    return avalanche(h);
}
constexpr uint64_t DIGEST_CHUNK = 8ull << 20;   // fixed, so the digest does not depend on the thread count
}  // namespace sym

extern "C" {
void sym_host_copy(void *dst, const void *src, uint64_t nbytes, int32_t threads) {
    // memcpy on a few threads: 4 MiB pieces handed out round-robin.  Staging a pageable input into page-locked memory is
    // bound by how many cores copy; torch's CPU copy does this with its intra-op threads, which a launcher may have
    // limited to one per process (torchrun sets OMP_NUM_THREADS=1).
    constexpr uint64_t PIECE = 4ull << 20;
    const uint64_t pieces = (nbytes + PIECE - 1) / PIECE;
    uint64_t nt = threads < 1 ? 1 : (uint64_t)threads;
    if (nt > pieces) nt = pieces;
    auto work = [&](uint64_t first, uint64_t step) {
        for (uint64_t i = first; i < pieces; i += step) {
            const uint64_t off = i * PIECE, len = nbytes - off < PIECE ? nbytes - off : PIECE;
            memcpy((unsigned char *)dst + off, (const unsigned char *)src + off, len);
        }
    };
    if (nt <= 1) {
        if (nbytes) memcpy(dst, src, nbytes);
        return;
    }
    std::vector<std::thread> pool;
    for (uint64_t t = 1; t < nt; ++t) pool.emplace_back(work, t, nt);
    work(0, nt);
    for (auto &t : pool) t.join();
}
uint64_t sym_host_digest64(const void *data, uint64_t nbytes, int32_t threads) {
    using namespace sym;
    const unsigned char *p = (const unsigned char *)data;
    const uint64_t chunks = (nbytes + DIGEST_CHUNK - 1) / DIGEST_CHUNK;
    if (chunks <= 1) return digest_span(p, nbytes, 0);
    std::vector<uint64_t> part(chunks);
    auto work = [&](uint64_t first, uint64_t step) {
        for (uint64_t c = first; c < chunks; c += step) {
            const uint64_t off = c * DIGEST_CHUNK, len = nbytes - off < DIGEST_CHUNK ? nbytes - off : DIGEST_CHUNK;
            part[c] = digest_span(p + off, len, c + 1);
        }
    };
    uint64_t nt = threads < 1 ? 1 : (uint64_t)threads;
    if (nt > chunks) nt = chunks;
    std::vector<std::thread> pool;
    for (uint64_t t = 1; t < nt; ++t) pool.emplace_back(work, t, nt);
    work(0, nt);
    for (auto &t : pool) t.join();
    return digest_span((const unsigned char *)part.data(), chunks * sizeof(uint64_t), nbytes);
}

int sym_version(void) { return 100; }
const char *sym_last_error(void) { return sym::g_err; }
int64_t sym_launch_count(void) { return sym::g_launches.load(std::memory_order_relaxed); }
int sym_compiled_arch(void) { return 100; }
}
