// reader_check.cpp -- CPU harness for fast_reader.h (tests/test_reader_cpu.py): frames the given files exactly like
// the drop-in front end does and prints, per file, "<n_reads> <n_bases> <fnv1a-64 over (length, bytes) of every read>".
//   reader_check <format 1|2> <block_bases> <block_reads> <trim> <file>...
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "fast_reader.h"

static int host_alloc(void **p, uint64_t n) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
static int host_free(void *p) { free(p); return 0; }

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: reader_check <format> <block_bases> <block_reads> <trim> <file>...\n"); return 2; }
    const int format = atoi(argv[1]);
    const uint64_t bb = strtoull(argv[2], nullptr, 10), br = strtoull(argv[3], nullptr, 10), trim = strtoull(argv[4], nullptr, 10);
    const bool nohash = getenv("READER_CHECK_NOHASH") != nullptr;      // throughput runs: count only
    std::vector<std::unique_ptr<dbgio::FileProducer> > prod;
    for (int i = 5; i < argc; i++)      // all files decode concurrently; consumed in order
        prod.emplace_back(new dbgio::FileProducer(argv[i], format, trim, bb, br, host_alloc, host_free));
    for (size_t f = 0; f < prod.size(); f++) {
        uint64_t n_reads = 0, n_bases = 0, h = 1469598103934665603ull;
        for (;;) {
            dbgio::ReadBlock *b = prod[f]->pop();
            if (prod[f]->failed()) { fprintf(stderr, "allocation failed\n"); return 1; }
            for (uint64_t i = 0; i < b->n_reads && !nohash; i++) {
                const uint64_t len = b->offs[i + 1] - b->offs[i];
                for (int k = 0; k < 8; k++) { h ^= (len >> (8 * k)) & 0xff; h *= 1099511628211ull; }
                for (uint64_t j = b->offs[i]; j < b->offs[i + 1]; j++) { h ^= (unsigned char)b->bases[j]; h *= 1099511628211ull; }
            }
            n_reads += b->n_reads; n_bases += b->n_bases;
            const bool last = b->last;
            prod[f]->recycle(b);
            if (last) break;
        }
        const std::string io = prod[f]->io_error();
        if (!io.empty()) fprintf(stderr, "input problem: %s\n", io.c_str());
        printf("%llu %llu %llu\n", (unsigned long long)n_reads, (unsigned long long)n_bases, (unsigned long long)h);
    }
    return 0;
}
