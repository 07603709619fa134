// TEST INFRASTRUCTURE ONLY -- harness around the reference's contig seed index.
//
// oracle/Makefile compiles this file together with /root/reference/link_scaffold/{kmerSet,map_func,seqKmer,gzstream}.cpp
// in place into oracle/_ref/ref_seed_driver.  It runs the UNMODIFIED read_contig_file, init_kmerset, chop_contig_to_kmerset
// and get_align_seed the way map_pair does (map_pair.cpp:97-125: contigs shorter than -l are blanked but keep their index,
// table of 3 x the contig length at load factor 0.5) and dumps
//   <out>.table : u64 size, count, max, conflict; then size x 16-byte nodes (zeros in empty slots: the reference leaves
//                 them uninitialised) and size/8+1 nul_flag bytes
//   <out>.seeds : per read (one sequence per line of <reads.txt>) six int32: contig_id_index, seed_contig_start,
//                 seed_contig_end, seed_read_start, seed_read_end, direct ('F','R','N')
//
//   ref_seed_driver <K> <min_ctg_len> <seed_kmer_num> <contigs.fa> <reads.txt> <out> [hash_size]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "map_func.h"

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: ref_seed_driver K min_ctg_len seed_kmer_num contigs.fa reads.txt out [hash_size]\n"); return 2; }
    KmerSize = atoi(argv[1]);
    MinCtgLen = atoi(argv[2]);
    SeedKmerNum = atoi(argv[3]);
    string contig_file = argv[4], reads_file = argv[5], out = argv[6];

    vector<string> contig_ids, contig_seqs;
    read_contig_file(contig_file, contig_ids, contig_seqs);
    uint64_t total_contig_len = 0;
    for (size_t i = 0; i < contig_seqs.size(); i++) {
        if (contig_seqs[i].size() >= (size_t)MinCtgLen) total_contig_len += contig_seqs[i].size();
        else contig_seqs[i] = "";
    }
    uint64_t hash_size = argc > 7 ? strtoull(argv[7], NULL, 10) : total_contig_len * 3;
    KmerSet *kset = init_kmerset(hash_size, 0.5);
    chop_contig_to_kmerset(kset, contig_seqs);

    {
        FILE *fp = fopen((out + ".table").c_str(), "wb");
        if (!fp) { perror("table"); return 1; }
        uint64_t hdr[4] = {kset->size, kset->count, kset->max, kset->count_conflict};
        fwrite(hdr, 8, 4, fp);
        KmerNode zero; memset(&zero, 0, sizeof zero);
        for (uint64_t i = 0; i < kset->size; i++)
            fwrite(is_entity_null(kset->nul_flag, i) ? &zero : kset->array + i, sizeof(KmerNode), 1, fp);
        fwrite(kset->nul_flag, 1, kset->size / 8 + 1, fp);
        fclose(fp);
    }
    {
        ifstream in(reads_file.c_str());
        FILE *fp = fopen((out + ".seeds").c_str(), "wb");
        if (!fp) { perror("seeds"); return 1; }
        string read;
        while (getline(in, read, '\n')) {
            int id = -1, cs = -1, ce = -1, rs = -1, re = -1; char d = 'N';
            if ((int)read.size() >= KmerSize + SeedKmerNum)                         // map_pair.cpp:284
                get_align_seed(kset, read, 1, read.size(), id, cs, ce, rs, re, d);
            int32_t rec[6] = {id, cs, ce, rs, re, (int32_t)d};
            fwrite(rec, 4, 6, fp);
        }
        fclose(fp);
    }
    return 0;
}
