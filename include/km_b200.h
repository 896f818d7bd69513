/* km_b200 -- C ABI of the B200-native find_mutation hot path of iric-soft/km.
 *
 * Drop-in boundary: the reference (pure Python, km 2.2.2) reaches its only native code,
 * the Jellyfish library, through four calls in km/utils/Jellyfish.py.  A km maintainer
 * would bind the functions below with ctypes (INTEGRATION.md shows the stub); this repo's
 * own km-compatible host layer (the modules of km_b200/utils) is that binding.
 *
 * Conventions: plain C types only.  Every function returning int returns 0 on success and
 * a negative KM_E_* code on failure; km_last_error() then holds a message (thread local).
 * Buffers named *_host are host pointers; *_dev are device pointers on the table's device.
 * k-mers are 2-bit packed uint64 (A0 C1 G2 T3, first base most significant), forward strand
 * as written in the sequence -- canonicalisation happens inside, exactly like
 * Jellyfish.query (km/utils/Jellyfish.py:47-53).  A missing k-mer counts 0, never an error
 * (km/tests/test_main.py:596-623).  One host thread per table handle.
 */
#ifndef KM_B200_H
#define KM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KM_E_ARG (-1)      /* bad argument */
#define KM_E_IO (-2)       /* file could not be read / unsupported .jf format */
#define KM_E_CUDA (-3)     /* CUDA runtime error (message has the detail) */
#define KM_E_FULL (-4)     /* table capacity exhausted */
#define KM_E_LIMIT (-5)    /* an internal capacity could not be satisfied after retries */
#define KM_E_NOGPU (-6)    /* no CUDA device: there is no CPU fallback */

typedef struct km_table km_table;
typedef struct km_result km_result;

const char* km_last_error(void);
int km_device_count(void);
const char* km_version(void);

/* ---- k-mer count table ------------------------------------------------------------
 * replaces jellyfish.QueryMerFile(filename) + jellyfish.MerDNA.k()  (Jellyfish.py:24-25)
 * and the JSON header parse for `canonical`                         (Jellyfish.py:28-45) */
int km_table_open_jf(const char* path, int device, km_table** out);
int km_table_create(int device, int k, int canonical, uint64_t capacity_keys, km_table** out);
/* The same with the table layout chosen explicitly: 0 = one 32-byte bucket per k-mer hash (one sector per
 * lookup), 1 = FAMILY LINES: 128-byte lines addressed by a canonical (k-1)-mer, holding the k-mers that start
 * or end with it; every k-mer is stored twice (~51 bytes per key instead of 32), and a k-mer's own count
 * plus its four successors -- what MutationFinder asks for together -- come from ONE line.  km_table_create
 * uses layout 0 unless the environment variable KM_TABLE_LINES is set. */
int km_table_create_layout(int device, int k, int canonical, uint64_t capacity_keys, int layout, km_table** out);
/* keys are canonical keys as stored in a .jf (`jellyfish count -C` records);
 * mode 0 keep existing, 1 overwrite, 2 add counts */
int km_table_insert(km_table* t, const uint64_t* keys_host, const uint32_t* counts_host, uint64_t n, int mode);
/* config-4 background: key_i = canonical(mix64(seed+(i+1)*G) & mask), count = f(key), i<n.
 * Generated and inserted on device (mode keep). */
int km_table_build_synthetic(km_table* t, uint64_t seed, uint64_t n_keys);
/* C19 / config 5: count canonical k-mers of `n_reads` reads (ASCII, concatenated, offsets
 * [n_reads+1]) on device; k-mers containing a non-ACGT letter are skipped. */
int km_table_count_reads(km_table* t, const char* reads_host, const int64_t* offsets_host, int64_t n_reads);
/* The same straight from a FASTA / FASTQ file, plain or .gz (zlib is looked up at run time), "-" = stdin: the
 * input side of `jellyfish count` (example/run_leucegene.sh:22).  min_qual_char > 0: bases whose FASTQ quality
 * character is below it count as N (`-Q`).  Reads and bases seen come back through the two counters. */
int km_table_count_file(km_table* t, const char* path, int min_qual_char, uint64_t* n_reads, uint64_t* n_bases);
/* The same from a BYTE STREAM: sequences separated by any byte outside ACGTacgt (a newline between reads), n_bytes
 * long; qual_host (may be NULL) = the FASTQ quality character of every byte, same length: with min_qual_char > 0 a
 * base whose quality is below it ends the k-mers that contain it (`jellyfish count -Q`), decided on the device.
 * The stream goes up in chunks through two pinned buffers, copies and kernels overlapping. */
int km_table_count_text(km_table* t, const char* text_host, const char* qual_host, uint64_t n_bytes, int min_qual_char);
/* drop entries with count < min_count (jellyfish count -L); returns remaining through *n_left */
int km_table_drop_below(km_table* t, uint32_t min_count, uint64_t* n_left);
/* `jellyfish dump`: every (canonical key, count) record of the table (of this shard in cohort mode), in no
 * particular order, into caller arrays of `cap` entries; *n_out = records in the table (may exceed cap). */
int km_table_export(km_table* t, uint64_t* keys_host, uint32_t* counts_host, uint64_t cap, uint64_t* n_out);
/* The table as a Jellyfish 2.x binary/sorted file -- what `jellyfish count -o` leaves on disk and
 * Jellyfish(filename) opens (km/utils/Jellyfish.py:23-45; layout: SURVEY.md Appendix A): records sorted by the
 * hash position M * key of the `matrix1` written into the header.  counter_len 0 = 4 bytes. */
int km_table_write_jf(km_table* t, const char* path, uint32_t counter_len);

/* Writes the NEIGHBOUR MASKS: for every stored k-mer, which of its 4 successors and 4 predecessors are stored too
 * (csrc/table.h).  km_find_* use them to answer "absent" for a successor without reading memory -- MutationFinder asks
 * for four successors per k-mer (Jellyfish.get_child, Jellyfish.py:61-66) and three of them are absent almost everywhere.
 * Called on demand by km_find_* after the table's content changed; call it yourself to keep that cost out of a timed
 * region.  No-op for cohort shards and the family-line layout. */
int km_table_link(km_table* t);

typedef struct km_table_info {
    int32_t k;
    int32_t canonical;
    int32_t device;
    int32_t reserved;      /* table layout: 0 sector buckets, 1 family lines */
    uint64_t n_keys;       /* distinct keys stored */
    uint64_t n_buckets;    /* 32-byte buckets */
    uint64_t bytes;        /* HBM held by the bucket array */
} km_table_info;
int km_table_get_info(km_table* t, km_table_info* info);
void km_table_close(km_table* t);

/* ---- cohort mode (BASELINE.json config 5; no counterpart in the reference): the table is
 * hash-sharded over the GPUs of one box, one process per GPU.  Every process creates ITS shard
 * (rank of n_shards <= 8, the same capacity on every rank), fills it -- km_table_insert /
 * km_table_build_synthetic / km_table_count_reads take the full key stream and keep only the keys this
 * shard owns -- and exports a POSIX file descriptor of its memory (CUDA virtual-memory API, 2 MiB pages).
 * The descriptors travel between the processes (SCM_RIGHTS; km_b200/cohort.py); once a process has
 * attached every other rank's descriptor, a probe of a remote key is a 32-byte peer load over NVLink from
 * inside the same kernels (km_query_batch, km_find_batch, ...).  Until then only owned keys may be asked. */
int km_table_create_shard(int device, int k, int canonical, uint64_t capacity_keys_per_shard, int rank, int n_shards,
                          km_table** out);
int km_table_shard_export_fd(km_table* t, int* fd);
int km_table_shard_attach_fd(km_table* t, int rank, int fd);      /* takes ownership of fd */
/* owner shard of each k-mer (forward-strand, packed): host arithmetic only, no GPU needed.  This is what
 * routes a query in the explicit all-to-all exchange (km_b200/cohort.py). */
int km_shard_owner(const uint64_t* kmers, uint64_t n, int k, int canonical, int n_shards, int32_t* owner);
/* Routed inserts (every peer attached first): with routing on, a key given to km_table_insert / km_table_count_* on
 * ANY rank is inserted into its OWNER's shard by system-scope atomics over NVLink -- extraction and exchange are one
 * kernel -- so every rank feeds its own part of the stream (its sample's reads).  With routing off (the default) a
 * shard keeps only the keys it owns and every rank must stream everything.  After routed inserts call
 * km_table_recount (behind a barrier): the keys of a shard were created by other ranks. */
int km_table_set_routing(km_table* t, int on);
int km_table_recount(km_table* t, uint64_t* n_keys);
/* The explicit exchange on the device: owner of each k-mer; a batch of queries grouped by owner with the permutation
 * that brings the answers back (counts_dev: 24 uint64 of scratch, [0..8) = queries per owner afterwards); the
 * inverse permutation of the answers.  All asynchronous on `cuda_stream` (NULL = the table's). */
int km_shard_owner_device(km_table* t, const uint64_t* kmers_dev, uint64_t n, int32_t* owner_dev, void* cuda_stream);
int km_route_partition(km_table* t, const uint64_t* kmers_dev, uint64_t n, uint64_t* sorted_dev, uint32_t* perm_dev,
                       uint64_t* counts_dev, void* cuda_stream);
int km_route_unpermute(km_table* t, const uint32_t* answers_dev, const uint32_t* perm_dev, uint64_t n, uint32_t* out_dev,
                       void* cuda_stream);

/* ---- lookups: Jellyfish.query (Jellyfish.py:47-53) ---------------------------------- */
int km_query_batch(km_table* t, const uint64_t* kmers_host, uint64_t n, uint32_t* counts_host);
int km_query_batch_device(km_table* t, const uint64_t* kmers_dev, uint64_t n, uint32_t* counts_dev, void* cuda_stream);
/* n k-mers of k ASCII letters each, back to back; a non-ACGT letter yields KM_E_ARG */
int km_query_ascii(km_table* t, const char* kmers_host, uint64_t n, uint32_t* counts_host);
/* Jellyfish.get_child (Jellyfish.py:55-72): for each k-mer the 4 neighbour counts (A,C,G,T)
 * and a 4-bit mask of those with count >= max(sum*ratio, count_floor) */
int km_get_child_batch(km_table* t, const uint64_t* kmers_host, uint64_t n, int forward, double ratio,
                       int64_t count_floor, uint32_t* child_counts_host /* 4n */, uint8_t* child_mask_host /* n */);

/* ---- find_mutation over a batch of targets -------------------------------------------
 * replaces the per-target loop of km/tools/find_mutation.py:47-58:
 *   MutationFinder(refpath, jf, steps, branchs, nodes)   km/utils/MutationFinder.py:87-134
 *   .graph_analysis()                                    :496-572 + km/utils/Graph.py
 *   .quantify_paths() / .quantify_clusters()             :575-811 + km/utils/PathQuant.py */
typedef struct km_find_params {
    double ratio;        /* -p, default 0.05 */
    int64_t count;       /* -c, default 5 */
    int32_t steps;       /* -s, default 500  (max_stack) */
    int32_t branchs;     /* -b, default 10   (max_break) */
    int32_t nodes;       /* -n, default 10000 (max_node) */
    int32_t extra_nodes; /* initial per-target capacity for explored novel nodes; 0 = default.
                            Targets that overflow are re-run with 8x until `nodes`-bounded */
    int32_t flags;       /* KM_FIND_* */
    int32_t reserved;
} km_find_params;
#define KM_FIND_NO_GRAPH 1   /* do not copy node arrays / index paths back (rows and text only) */
#define KM_FIND_NO_REFINE_JUMP 2   /* iterate PathQuant.refine_coef (PathQuant.py:120-142) literally from start to end: no
                                      closed-form jump across its long linear stretches (A/B switch; the environment
                                      variable KM_NO_REFINE_JUMP does the same) */

/* per-target status bits */
#define KM_ST_BAD_BASE 1
#define KM_ST_DUP_KMER 2
#define KM_ST_NODE_OVERFLOW 4
#define KM_ST_NODE_LIMIT 8
#define KM_ST_TOUCHED_LIMIT 16
#define KM_ST_PATH_OVERFLOW 32
#define KM_ST_TOO_SHORT 64
#define KM_ST_TOO_MANY_COLS 128
#define KM_ST_SOLVER_WATCHDOG 256
#define KM_ST_NAME_MISMATCH 512

/* One output row (km/utils/PathQuant.py:10-49) in numeric form. */
typedef struct km_row {
    int32_t target;
    int32_t kind;               /* 0 vs_ref, 1 cluster */
    int32_t type;               /* 0 Reference 1 Substitution 2 ITD 3 Indel 4 Insertion 5 Deletion */
    int32_t name_start;
    int32_t name_end;
    int32_t path_id;
    int32_t var_begin, var_end;
    int32_t ref_begin, ref_end;
    int32_t del_begin, del_len;
    int32_t ins_begin, ins_len;
    int32_t start_off;
    int32_t cluster_id, cluster_n;
    int32_t n_iter;
    int64_t min_cov;
    double rvaf, expr, ref_rvaf, ref_expr;
} km_row;

/* seqs_host: the targets' sequences concatenated (ASCII, already upper-cased like
 * common.file_2_seq, common.py:43); offsets_host[n_targets+1]. */
int km_find_batch(km_table* t, const char* seqs_host, const int64_t* offsets_host, int32_t n_targets,
                  const km_find_params* params, km_result** out);

/* The same pipeline in three steps, for callers that keep a batch resident in HBM and launch it
 * repeatedly (bench.py): create = size + upload inputs, launch = the two kernels on a stream
 * (asynchronous, NULL = the table's stream), fetch = synchronise + copy results back (re-running
 * with larger capacities if any target overflowed). */
typedef struct km_plan km_plan;
int km_find_plan_create(km_table* t, const char* seqs_host, const int64_t* offsets_host, int32_t n_targets,
                        const km_find_params* params, km_plan** out);
int km_find_plan_launch(km_plan* p, void* cuda_stream);
/* device time of the most recent launch: memsets + walk kernel, and graph kernel (CUDA events) */
int km_find_plan_last_ms(km_plan* p, float* walk_ms, float* graph_ms);
/* the same split three ways: out3[0] memsets + reference-probe kernel, [1] walk kernels, [2] graph kernels */
int km_find_plan_kernel_ms(km_plan* p, float* out3);
int km_find_plan_fetch(km_plan* p, int want_graph, km_result** out);
void km_find_plan_free(km_plan* p);

/* flat views into a result (valid until km_result_free) */
typedef struct km_result_view {
    int32_t n_targets;
    int32_t n_paths;
    int32_t n_rows;
    int32_t k;
    const uint32_t* status;        /* [n_targets] KM_ST_* bits */
    const int32_t* n_nodes;        /* [n_targets] graph nodes incl. BigBang/BigCrunch */
    const int64_t* node_off;       /* [n_targets+1] into node_kmer/node_count */
    const uint64_t* node_kmer;     /* canonical numbering: reference k-mers, then novel ascending */
    const uint32_t* node_count;
    const int32_t* path_first;     /* [n_targets] */
    const int32_t* path_count;     /* [n_targets] */
    const int64_t* path_off;       /* [n_paths] into path_pool */
    const int32_t* path_len;       /* [n_paths] */
    const int32_t* path_pool;
    const int32_t* row_first;      /* [n_targets] */
    const int32_t* row_count;      /* [n_targets] */
    const km_row* rows;            /* [n_rows] quantify_paths rows then quantify_clusters rows per target */
    const uint64_t* lookups;       /* [n_targets] table lookups issued by the walk */
    /* device time of the last run, milliseconds (CUDA events on the launch stream) */
    float ms_h2d, ms_walk, ms_graph, ms_d2h, ms_total;
    int32_t n_launches;            /* kernels launched for this result */
    int32_t n_retries;
    int32_t has_graph;             /* 0: node_kmer/node_count/path_pool were not copied back */
    int32_t reserved;              /* measurement: targets whose graph was a "simple bubble" (closed-form paths, csrc/graph.h) */
    uint64_t bytes_h2d, bytes_d2h; /* bytes copied host->device / device->host for this result */
} km_result_view;
int km_result_get(const km_result* r, km_result_view* view);
void km_result_free(km_result* r);

/* Formats the rows of one target exactly like km find_mutation prints them (str(Path),
 * PathQuant.py:37-49, sorted as MutationFinder.get_paths, :825-829).  Returns the number
 * of bytes written (excluding the NUL), or the required size if buf is too small. */
int64_t km_result_format_target(const km_result* r, int32_t target, const char* db_name, const char* query_name,
                                char* buf, int64_t buf_len);
/* All targets in order, names_host = query names concatenated, name_off[n_targets+1].  Formats
 * on `threads` host threads (0 = hardware concurrency). */
int64_t km_result_format_all(const km_result* r, const char* db_name, const char* names_host, const int64_t* name_off,
                             int32_t threads, char* buf, int64_t buf_len);

/* The same text without a copy: *text points at the result's own buffer (valid until km_result_free or a
 * call with other names); returns its length. */
int64_t km_result_text(const km_result* r, const char* db_name, const char* names_host, const int64_t* name_off,
                       int32_t threads, const char** text);

/* `km find_mutation` for a whole batch as ONE call: host buffers in (sequences + offsets, query names +
 * offsets), the sorted TSV text out (read it with km_result_text, same db_name / names; km_result_get
 * gives the per-target status).  The batch runs as n_sub sub-batches in flight at once (0 = choose), so
 * copies and kernels overlap; the text itself is formatted ON THE DEVICE (csrc/format.h) and every sub-batch's
 * share comes back with one copy, straight into its place in the result's pinned buffer.  Byte for byte the
 * text of km_find_batch + km_result_format_all (the host formatter, which a sub-batch falls back to when the
 * device declines: text capacity, a magnitude >= 2^52; KM_HOST_FORMAT=1 forces it). */
int km_find_text(km_table* t, const char* seqs_host, const int64_t* offsets, int32_t n_targets, const km_find_params* params,
                 const char* db_name, const char* names_host, const int64_t* name_off, int32_t n_sub, km_result** out);

/* ---- test hooks for the host-side text path (no GPU needed) ------------------------------ */
/* "%.{prec}f" as the formatter prints it (prec 0..3), NUL-terminated into buf64; returns the length */
int km_debug_format_fixed(double v, int prec, char* buf64);
/* natural-sort comparison of two strings (common.natsortkey, common.py:95-116): -1, 0, 1 */
int km_debug_nat_cmp(const char* a, const char* b);

/* ---- measurement helpers (bench.py) --------------------------------------------------- */
/* per-phase SM cycles of the graph pass (only in a -DKM_PHASE_TIMERS build; tools/phase_times.py) */
int km_debug_phase_cycles(unsigned long long* out64, int reset);
/* the same for the walk kernels (their counters live in their own translation unit) */
int km_debug_walk_cycles(unsigned long long* out64, int reset);
/* graph-pass SM cycles of targets 0..n-1 in the last launch (KM_PHASE_TIMERS builds only) */
int km_debug_target_cycles(unsigned int* out, int n);   /* 64 counters */
/* wall-clock extent (%globaltimer ns) of every kernel of the graph phase in the launches since the last reset
 * (-DKM_TIMELINE or -DKM_PHASE_TIMERS builds only; tools/slow_targets.py): out32 = 16 x (earliest CTA start, latest CTA
 * end); 0 scheduler, 1 / 2 / 3 CTA-per-target passes (256 / 512 nodes / general), 4 / 5 bubble passes, 8.. the first
 * eight targets the CTA-per-target passes took */
int km_debug_timeline(unsigned long long* out32, int reset);
/* random 32-byte-sector gather over `bytes` of HBM: the ceiling for hash probes */
int km_bench_random_gather(int device, uint64_t bytes, uint64_t n_loads, int iters, float* best_ms);
/* device-resident counting benchmark: n_reads reads of read_len bases drawn from a pseudo-random genome of `genome` bases
 * are generated on the device and counted iters + 1 times (km_count_text_kernel alone, CUDA events): best_ms2[0] = best
 * pass over existing keys, best_ms2[1] = the first pass, which creates them */
int km_bench_count(km_table* t, uint64_t n_reads, int read_len, uint64_t genome, uint64_t seed, int iters, float* best_ms2,
                   uint64_t* n_kmers);
/* the config-4 query mix (50 % background keys on a random strand, 50 % random k-mers) into a caller's device buffer */
int km_bench_make_queries(km_table* t, uint64_t* queries_dev, uint64_t n, uint64_t table_seed, uint64_t table_n,
                          uint64_t query_seed, void* cuda_stream);
/* device-resident lookup benchmark: n queries (config-4 mix) generated on device, timed `iters` times */
int km_bench_lookup(km_table* t, uint64_t table_seed, uint64_t table_n, uint64_t n_queries, uint64_t query_seed,
                    int iters, float* best_ms, float* mean_ms, uint64_t* n_hits);

#ifdef __cplusplus
}
#endif
#endif /* KM_B200_H */
