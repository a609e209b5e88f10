// Compile check, host only: a caller written against the REFERENCE's spelling of the classify path -- Database<khash_t(c)>,
// ClassifierGeneric<score::Lex>(db.db_, db.s_, db.k_, ...), khash_t(p) *build_parent_map, process_dataset, classify_seqs with a
// ks::string and a ForPool, kh_destroy(p, ...) -- the calls bin/bonsai.cpp:149-160 and include/bonsai/classifier.h:269-337 make,
// against include/bonsai_b200/bonsai.hpp. Linked with tests/host/abi_stub.cpp by tests/test_cli_cpu.py; classifies nothing real.
#include "../../include/bonsai_b200/bonsai.hpp"
using namespace bns;
int main(int argc, char **argv) {
    if(argc < 4) return 2;
    const int num_threads = 2, emit_all = 1, emit_fastq = 0, emit_kraken = 1, chunk_size = 1 << 16, per_set = 32;
    const bool canonicalize = true;
    try {
        Database<khash_t(c)> db(argv[1]);
        ClassifierGeneric<score::Lex> c(db.db_, db.s_, db.k_, db.k_, num_threads, emit_all, emit_fastq, emit_kraken, canonicalize);
        khash_t(p) *taxmap(build_parent_map(argv[2]));
        process_dataset(c, taxmap, argv[3], argc > 4 ? argv[4] : nullptr, stdout, chunk_size, per_set);
        // the batch call with the reference's own parameter list
        std::vector<bseq1_t> bs(2);
        bs[0].name = "a"; bs[0].seq = std::string(60, 'A'); bs[0].l_seq = 60;
        bs[1].name = "b"; bs[1].seq = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"; bs[1].l_seq = (int)bs[1].seq.size();
        ks::string cks(256u);
        ForPool pool(c.nt_);
        classify_seqs(c, taxmap, bs.data(), cks, 2, per_set, 0, pool);
        cks.write(fileno(stdout));
        cks.clear();
        kh_destroy(p, taxmap);
    } catch(const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
