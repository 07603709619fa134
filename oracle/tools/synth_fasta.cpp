// TEST / BENCH INFRASTRUCTURE -- stand-alone writer of the synthetic read sets as one-line FASTA, so
// that `bench.py --impl reference` and the cpu_baseline leg can hand the reference program its input without loading
// anything from the product library.  Same generator as libdbgb200 (dbg_assembly_b200/csrc/synth_core.h: counter based,
// byte-identical on host and device).
//
//   synth_fasta <seed> <genome_len> <read_len> <insert> <err_per_2^24> <n_per_2^24> <first_read> <n_reads> <out.fa> [raw]
//
// "raw": bases only, no headers / newlines (for the oracle port, which takes arrays).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../dbg_assembly_b200/csrc/synth_core.h"

int main(int argc, char **argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: synth_fasta seed genome_len read_len insert err_per_2p24 n_per_2p24 first_read n_reads out.fa [raw]\n");
        return 2;
    }
    synth_params_t p;
    p.seed = strtoull(argv[1], nullptr, 10); p.genome_len = strtoull(argv[2], nullptr, 10);
    p.read_len = (uint32_t)strtoul(argv[3], nullptr, 10); p.insert = (uint32_t)strtoul(argv[4], nullptr, 10);
    p.err_per_2p24 = (uint32_t)strtoul(argv[5], nullptr, 10); p.n_per_2p24 = (uint32_t)strtoul(argv[6], nullptr, 10);
    const uint64_t first = strtoull(argv[7], nullptr, 10), n = strtoull(argv[8], nullptr, 10);
    const bool raw = argc > 10 && strcmp(argv[10], "raw") == 0;
    const uint32_t ins = p.insert < p.read_len ? p.read_len : p.insert;
    if (p.read_len == 0 || p.read_len > 65535 || p.genome_len < ins) { fprintf(stderr, "synth_fasta: bad parameters\n"); return 2; }
    FILE *fp = fopen(argv[9], "wb");
    if (!fp) { perror(argv[9]); return 1; }
    // fixed-width records ('>' + 11 digits + '\n' + bases + '\n'): threads fill disjoint slices of a chunk buffer
    const uint64_t rec = raw ? p.read_len : (uint64_t)p.read_len + 14;
    const uint64_t CHUNK = 1u << 16;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    std::vector<char> buf(CHUNK * rec);
    for (uint64_t r0 = 0; r0 < n; r0 += CHUNK) {
        const uint64_t m = n - r0 < CHUNK ? n - r0 : CHUNK;
        auto work = [&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; i++) {
                char *q = buf.data() + i * rec;
                if (!raw) { char h[16]; snprintf(h, sizeof(h), ">%011llu\n", (unsigned long long)(first + r0 + i)); memcpy(q, h, 13); q += 13; }
                for (uint32_t t = 0; t < p.read_len; t++) q[t] = synth::synth_base(p, first + r0 + i, t);
                if (!raw) q[p.read_len] = '\n';
            }
        };
        std::vector<std::thread> th;
        const uint64_t per = (m + nt - 1) / nt;
        for (unsigned k = 1; k < nt; k++) { const uint64_t a = k * per, b = a + per < m ? a + per : m; if (a < b) th.emplace_back(work, a, b); }
        work(0, per < m ? per : m);
        for (auto &t : th) t.join();
        if (fwrite(buf.data(), rec, m, fp) != m) { perror("write"); return 1; }
    }
    if (fclose(fp) != 0) { perror("close"); return 1; }
    return 0;
}
