// CPU test of the streamer's host-side machinery (simlod_b200/csrc/loader_pool.h), built by
// tests/test_loader_pool_cpu.py with -fsanitize=thread and -fsanitize=address,undefined.
//  * LoaderPool: every generation runs exactly `n` workers, each with a distinct index and its own staging buffer;
//    wait() returns only when all of them are done; the pool grows on demand and can be reused many times.
//  * the piece protocol of simlod_insert_simlod_file in miniature: workers take 64 KB pieces of a "file" in order from
//    an atomic cursor, stage them and stream them into slots of a destination; the result must equal the source.
#include <atomic>
#include <cstdio>
#include <numeric>
#include "../../simlod_b200/csrc/loader_pool.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c); fails++; } } while (0)

int main() {
    // ---- copyStreaming: every 16-byte multiple up to a few KB, arbitrary 16-aligned offsets
    {
        const size_t N = 8192;
        char* src = (char*)aligned_alloc(64, N);
        char* dst = (char*)aligned_alloc(64, N + 64);
        for (size_t i = 0; i < N; i++) src[i] = (char)(i * 131 + 7);
        for (size_t bytes = 0; bytes <= 4096; bytes += 16)
            for (size_t off : {size_t(0), size_t(16), size_t(48)}) {
                memset(dst, 0xEE, N + 64);
                copyStreaming(dst + off, src + off, bytes);
                CHECK(memcmp(dst + off, src + off, bytes) == 0);
                CHECK((unsigned char)dst[off + bytes] == 0xEE);
                if (off) CHECK((unsigned char)dst[off - 1] == 0xEE);
            }
        free(src); free(dst);
    }
    // ---- pool generations
    {
        LoaderPool pool;
        const int sizes[] = {1, 4, 16, 8, 32, 3, 32, 1};
        for (int rep = 0; rep < 40; rep++)
            for (int n : sizes) {
                std::atomic<int> ran{0};
                std::vector<std::atomic<int>> seen(64);
                for (auto& s : seen) s.store(0);
                pool.run(n, [&](int worker) {
                    seen[worker].fetch_add(1);
                    pool.bounce[worker][0] = (char)worker;              // own staging buffer, no sharing
                    pool.bounce[worker][LoaderPool::BOUNCE_BYTES - 1] = (char)worker;
                    ran.fetch_add(1);
                });
                pool.wait();
                CHECK(ran.load() == n);
                for (int w = 0; w < 64; w++) CHECK(seen[w].load() == (w < n ? 1 : 0));
            }
        CHECK((int)pool.threads.size() == 32);
    }
    // ---- the piece protocol in miniature (pieces in file order, staged, streamed into their slot)
    {
        LoaderPool pool;
        const uint64_t PIECE = 64 << 10, TOTAL = (5u << 20) + 4096 + 16;     // ragged tail, multiple of 16
        std::vector<char> file(TOTAL);
        for (uint64_t i = 0; i < TOTAL; i++) file[i] = (char)((i * 2654435761u) >> 13);
        char* dst = (char*)aligned_alloc(64, (TOTAL + 64 + 63) / 64 * 64);
        for (int threads : {1, 3, 16}) {
            memset(dst, 0, TOTAL + 64);
            std::atomic<uint64_t> next{0};
            std::atomic<uint64_t> piecesDone{0};
            const uint64_t pieces = (TOTAL + PIECE - 1) / PIECE;
            pool.run(threads, [&](int worker) {
                for (;;) {
                    const uint64_t k = next.fetch_add(1);
                    if (k >= pieces) break;
                    uint64_t at = k * PIECE, bytes = std::min<uint64_t>(PIECE, TOTAL - at);
                    char* stage = pool.bounce[worker];
                    memcpy(stage, file.data() + at, bytes);            // stands in for pread into the staging buffer
                    copyStreaming(dst + at, stage, bytes);
                    piecesDone.fetch_add(1);
                }
            });
            pool.wait();
            CHECK(piecesDone.load() == pieces);
            CHECK(memcmp(dst, file.data(), TOTAL) == 0);
        }
        free(dst);
    }
    if (fails) { fprintf(stderr, "%d checks failed\n", fails); return 1; }
    printf("loader pool ok\n");
    return 0;
}
