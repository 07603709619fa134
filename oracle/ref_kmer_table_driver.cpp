// TEST INFRASTRUCTURE ONLY -- harness around the reference's in-tree K-mer table builder.
//
// oracle/Makefile compiles this file together with /root/reference/correct_error/{simulate_lowfreq_kmer,seqKmer,
// gzstream}.cpp in place (the tool's own main() renamed with -Dmain=sim_tool_main) into
// oracle/_ref/ref_kmer_table_driver.  It calls the UNMODIFIED construct_ref_kmer_table
// (correct_error/simulate_lowfreq_kmer.cpp:189-260: one bit per K-mer of the genome AND of its reverse complement, MSB
// first, array of total/8+1 bytes) and writes the raw bit table to a file, so that tests can compare the B200 table
// (kfreq_export bits=1, reverse-complement bits OR-ed in the way correct_error's loader does) bit for bit with it.
//
//   ref_kmer_table_driver <K> <genome.fa> <out.bits>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <inttypes.h>

uint8_t *construct_ref_kmer_table(std::string &genome_seq_file, int KmerSize, uint64_t &total);
extern int KmerSize;

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ref_kmer_table_driver K genome.fa out.bits\n"); return 2; }
    KmerSize = atoi(argv[1]);
    std::string fa = argv[2];
    uint64_t total = 0;
    uint8_t *bits = construct_ref_kmer_table(fa, KmerSize, total);
    FILE *fp = fopen(argv[3], "wb");
    if (!fp) { perror(argv[3]); return 1; }
    const uint64_t nbytes = total / 8 + 1;
    if (fwrite(bits, 1, nbytes, fp) != nbytes) { perror("write"); return 1; }
    fclose(fp);
    return 0;
}
