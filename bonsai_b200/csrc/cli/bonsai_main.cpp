// bonsai (B200) -- the `classify` subcommand of bin/bonsai.cpp:107-163 with its flag letters, over libbonsai_b200.so.
// Also: `build` (a working replacement for the reference's broken phase2/build, SURVEY App. B-1/B-6: GPU encoder +
// update_lca_map merge -> the DB file layout the reference intends), `dbwrite` / `dbcheck` (DB file tooling).
#include <getopt.h>


#include "../../../include/bonsai_b200/bonsai.hpp"

using namespace bns;

static int classify_usage(const char *ex) {
    std::fprintf(stderr,
                 "Usage:\n%s classify <opts> <dbpath> <tax_path> <inr1.fq> <inr2.fq>\nFlags:\n"
                 "-o:\tRedirect output to path instead of stdout.\n-c:\tSet chunk size [1048576 bases]\n"
                 "-a:\tEmit all records, not just classified.\n-p:\tSet number of threads. Default: 1.\n"
                 "-k:\tEmit kraken-style output.\n-K:\tDo not emit kraken-style output.\n-f:\tEmit fastq-style output.\n"
                 "-F:\tDo not emit fastq-style output.\n-C:\tDo not canonicalize.\n-S:\tSet records per worker set (ignored: one GPU call per chunk)\n"
                 "--gpus N:\tShard the reads over N GPUs (database replicated once over NVLink). Default: 1.\n",
                 ex);
    return EXIT_FAILURE;
}

// classify_main, bin/bonsai.cpp:107-163
static int classify_main(int argc, char *argv[]) {
    int co, num_threads(1), emit_kraken(1), emit_fastq(0), emit_all(0), chunk_size(1 << 20), per_set(32), n_gpus(1);
    bool canonicalize(true);
    std::FILE *ofp(stdout);
    if(argc < 4) return classify_usage(argv[0]);
    static const struct option long_opts[] = {{"gpus", required_argument, nullptr, 1000}, {nullptr, 0, nullptr, 0}};
    while((co = getopt_long(argc, argv, "Cc:p:o:S:afFkKh?", long_opts, nullptr)) >= 0) {
        switch(co) {
            case 1000: n_gpus = std::atoi(optarg); break;
            case 'h': case '?': return classify_usage(argv[0]);
            case 'a': emit_all = 1; break;
            case 'C': canonicalize = false; break;
            case 'c': chunk_size = std::atoi(optarg); break;
            case 'F': emit_fastq = 0; break;
            case 'f': emit_fastq = 1; break;
            case 'K': emit_kraken = 0; break;
            case 'k': emit_kraken = 1; break;
            case 'p': num_threads = std::atoi(optarg); break;
            case 'S': per_set = std::atoi(optarg); break;
            case 'o': ofp = std::fopen(optarg, "w"); if(!ofp) { std::fprintf(stderr, "Could not open %s\n", optarg); return EXIT_FAILURE; } break;
        }
    }
    if(argc - optind < 3) return classify_usage(argv[0]);
    try {
        const bool verbose = std::getenv("BNS_B200_VERBOSE") != nullptr;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        Database db(argv[optind]);
        const double t1 = now();
        // bin/bonsai.cpp:152: always score::Lex with window = k, whatever the DB was minimised with
        Classifier c(db, db.s_, (u8)db.k_, (u16)db.k_, num_threads, emit_all, emit_fastq, emit_kraken, canonicalize);
        if(n_gpus > 1) c.set_gpus(n_gpus);       // replicas are made once the taxonomy is loaded (process_dataset)
        const double t2 = now();
        std::unique_ptr<TaxMap> taxmap(build_parent_map(argv[optind + 1]));
        const char *fq2 = (argc - optind >= 4) ? argv[optind + 3] : nullptr;
        const double t3 = now();
        c.set_exit_follows(true);
        process_dataset(c, taxmap.get(), argv[optind + 2], fq2, ofp, (unsigned)chunk_size, (unsigned)per_set);
        if(verbose)
            std::fprintf(stderr, "[bonsai classify] database file %.2f s, classifier (device open + table load) %.2f s, taxonomy %.2f s, "
                         "dataset %.2f s\n", t1 - t0, t2 - t1, t3 - t2, now() - t3);
        std::fprintf(stderr, "Successfully finished classify_main. classified %" PRIu64 ", unclassified %" PRIu64 "\n",
                     c.n_classified(), c.n_unclassified());
        // Everything is written: the process ends here instead of unpinning buffers, freeing device memory and tearing the CUDA
        // context down object by object (0.2 - 0.5 s of a run whose work took as long)
        if(ofp != stdout) std::fclose(ofp);
        std::fflush(nullptr);
        std::_Exit(EXIT_SUCCESS);
    } catch(const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
    }
    if(ofp != stdout) std::fclose(ofp);
    return EXIT_SUCCESS;
}

static int build_usage(const char *ex) {
    std::fprintf(stderr,
                 "Usage:\n%s build <opts> <dbpath> <tax_path> <taxid=genome.fa[.gz]> ...\nFlags:\n-k:\tk-mer length [31]\n"
                 "-w:\twindow size [k]\n-s:\tspacing string, e.g. 1x3,0x5 [unspaced]\n-e:\tminimise by entropy instead of Lex\n"
                 "-C:\tDo not canonicalize\n-z:\tgzip the database\n", ex);
    return EXIT_FAILURE;
}

static int build_main(int argc, char *argv[]) {
    int co, k(31), w(-1);
    bool canon(true), entropy(false), gz(false);
    std::string spacing;
    while((co = getopt(argc, argv, "k:w:s:eCzh?")) >= 0) {
        switch(co) {
            case 'k': k = std::atoi(optarg); break;
            case 'w': w = std::atoi(optarg); break;
            case 's': spacing = optarg; break;
            case 'e': entropy = true; break;
            case 'C': canon = false; break;
            case 'z': gz = true; break;
            default: return build_usage(argv[0]);
        }
    }
    if(argc - optind < 3) return build_usage(argv[0]);
    try {
        const spvec_t gaps = parse_spacing(spacing.c_str(), k);
        Spacer sp(k, w < 0 ? k : w, gaps);
        std::unique_ptr<TaxMap> tm(build_parent_map(argv[optind + 1]));
        std::vector<std::pair<tax_t, std::string>> genomes;
        for(int i = optind + 2; i < argc; ++i) {
            const char *eq = std::strchr(argv[i], '=');
            if(!eq) { std::fprintf(stderr, "expected taxid=path, got %s\n", argv[i]); return EXIT_FAILURE; }
            genomes.emplace_back((tax_t)std::atoi(argv[i]), std::string(eq + 1));
        }
        // fill_set_genome + update_lca_map (feature_min.h:68-83,205-228) on the device: every genome's records go through
        // the encoder's record overloads and an insert-or-LCA-merge into the device table, in one kernel per genome.
        auto h = detail::open_handle(sp, entropy ? BNS_SCORE_ENTROPY : BNS_SCORE_LEX, canon && sp.unspaced(), BNS_API_PATH);
        detail::check(h->h, bns_b200_load_taxonomy(h->h, tm->child.data(), tm->parent.data(), tm->size()), "bns_b200_load_taxonomy");
        std::vector<std::string> bases(genomes.size());
        std::vector<std::vector<u64>> offs(genomes.size());
        std::vector<tax_t> taxids;
        u64 bound = 1024;
        for(size_t g = 0; g < genomes.size(); ++g) {
            detail::KSeq ks(genomes[g].second.c_str());
            offs[g].push_back(0);
            while(ks.read() >= 0) { bases[g] += ks.seq; offs[g].push_back(bases[g].size()); }
            taxids.push_back(genomes[g].first);
            bound += bases[g].size();
        }
        for(;; bound *= 2) {
            detail::check(h->h, bns_b200_build_begin(h->h, bound, taxids.data(), (u32)taxids.size()), "bns_b200_build_begin");
            for(size_t g = 0; g < genomes.size(); ++g)
                detail::check(h->h, bns_b200_build_add_genome(h->h, bases[g].data(), offs[g].data(), offs[g].size() - 1, taxids[g]),
                              "bns_b200_build_add_genome");
            const int rc = bns_b200_build_finish(h->h);
            if(rc == BNS_OK) break;
            if(rc != BNS_E_CAPACITY) detail::check(h->h, rc, "bns_b200_build_finish");
        }
        u64 n = 0;
        detail::check(h->h, bns_b200_table_dump(h->h, nullptr, nullptr, 0, &n), "bns_b200_table_dump");
        std::vector<u64> keys(n); std::vector<u32> vals(n);
        if(n) detail::check(h->h, bns_b200_table_dump(h->h, keys.data(), vals.data(), n, &n), "bns_b200_table_dump");
        {   // deterministic file: insert in key order
            std::vector<std::pair<u64, u32>> kv(n);
            for(u64 i = 0; i < n; ++i) kv[i] = {keys[i], vals[i]};
            std::sort(kv.begin(), kv.end());
            for(u64 i = 0; i < n; ++i) { keys[i] = kv[i].first; vals[i] = kv[i].second; }
        }
        std::fprintf(stderr, "[build] %zu genomes -> %" PRIu64 " distinct k-mers\n", genomes.size(), n);
        Database db;
        db.assign(k, sp.w_, gaps, keys.data(), vals.data(), keys.size());
        db.write(argv[optind], gz);
        std::fprintf(stderr, "[build] wrote %s: k=%d w=%u keys=%zu buckets=%" PRIu64 "\n", argv[optind], k, sp.w_, keys.size(), db.n_buckets);
    } catch(const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}

// dbwrite <out.db> <k> <w> <pairs.bin>: pairs.bin = u64 n, u64 keys[n], u32 vals[n]   (host only)
static int dbwrite_main(int argc, char *argv[]) {
    if(argc < 6) { std::fprintf(stderr, "Usage: %s dbwrite <out.db> <k> <w> <pairs.bin> [gz]\n", argv[0]); return EXIT_FAILURE; }
    try {
        std::FILE *fp = std::fopen(argv[5], "rb");
        if(!fp) BNS_RUNTIME_ERROR("cannot open pairs file");
        u64 n = 0;
        if(std::fread(&n, 8, 1, fp) != 1) BNS_RUNTIME_ERROR("short pairs file");
        std::vector<u64> keys(n); std::vector<u32> vals(n);
        if(n && (std::fread(keys.data(), 8, n, fp) != n || std::fread(vals.data(), 4, n, fp) != n)) BNS_RUNTIME_ERROR("short pairs file");
        std::fclose(fp);
        Database db;
        db.assign(std::atoi(argv[3]), std::atoi(argv[4]), spvec_t{}, keys.data(), vals.data(), n);
        db.write(argv[2], argc > 6);
    } catch(const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return EXIT_FAILURE; }
    return EXIT_SUCCESS;
}

// dbcheck <db>: header + xor/sum digests of the occupied (key, value) pairs   (host only)
static int dbcheck_main(int argc, char *argv[]) {
    if(argc < 3) { std::fprintf(stderr, "Usage: %s dbcheck <db>\n", argv[0]); return EXIT_FAILURE; }
    try {
        Database db(argv[2]);
        u64 n = 0, kx = 0, ks = 0, vs = 0;
        for(u64 i = 0; i < db.n_buckets; ++i)
            if(db.exists(i)) { ++n; kx ^= db.keys[i]; ks += db.keys[i]; vs += db.vals[i]; }
        std::printf("k=%u w=%u n_buckets=%" PRIu64 " size=%" PRIu64 " occupied=%" PRIu64 " key_xor=%016" PRIx64 " key_sum=%016" PRIx64 " val_sum=%" PRIu64 "\n",
                    db.k_, db.w_, db.n_buckets, db.size, n, kx, ks, vs);
    } catch(const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return EXIT_FAILURE; }
    return EXIT_SUCCESS;
}

// hist_main, bin/bonsai.cpp:351-374: how many k-mers each taxid owns, ascending by count (ties by taxid)   (host only)
static int hist_main(int argc, char *argv[]) {
    if(argc < 3 || !std::strcmp(argv[2], "-h") || !std::strcmp(argv[2], "--help")) {
        std::fprintf(stderr, "Produces a histogram of how many kmers have been assigned to all taxids\n"
                             "Usage: bonsai %s <database.db> [outfile (omit to emit to stdout)]\n", argv[1]);
        return EXIT_FAILURE;
    }
    try {
        Database db(argv[2]);
        std::vector<tax_t> v;
        for(u64 i = 0; i < db.n_buckets; ++i) if(db.exists(i)) v.push_back(db.vals[i]);
        std::sort(v.begin(), v.end());
        std::vector<std::pair<u32, tax_t>> structs;
        for(size_t i = 0; i < v.size();) {
            size_t j = i;
            while(j < v.size() && v[j] == v[i]) ++j;
            structs.emplace_back((u32)(j - i), v[i]);
            i = j;
        }
        std::sort(structs.begin(), structs.end());
        std::FILE *ofp = argc > 3 ? std::fopen(argv[3], "w") : stdout;
        if(!ofp) BNS_RUNTIME_ERROR("cannot open output file");
        std::fputs("Name\tCount\n", ofp);
        for(const auto &e : structs) std::fprintf(ofp, "%u\t%u\n", e.second, e.first);
        if(ofp != stdout) std::fclose(ofp);
    } catch(const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return EXIT_FAILURE; }
    return EXIT_SUCCESS;
}

static int usage(const char *ex) {
    std::fprintf(stderr, "Usage: %s <subcommand> [options...]. Use %s <subcommand> for more options.\n"
                         "Subcommands:\nclassify\nbuild\nhist\ndbwrite\ndbcheck\n", ex, ex);
    return EXIT_FAILURE;
}

int main(int argc, char *argv[]) {                          // bin/bonsai.cpp:521-540
    if(argc < 2) return usage(argv[0]);
    const std::string cmd(argv[1]);
    if(cmd == "classify") return classify_main(argc - 1, argv + 1);
    if(cmd == "build" || cmd == "phase2" || cmd == "p2") return build_main(argc - 1, argv + 1);
    if(cmd == "hist") return hist_main(argc, argv);
    if(cmd == "dbwrite") return dbwrite_main(argc, argv);
    if(cmd == "dbcheck") return dbcheck_main(argc, argv);
    return usage(argv[0]);
}
