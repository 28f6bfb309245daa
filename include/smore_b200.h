/* smore_b200 -- C ABI of the B200-native sampled-SGD backend for SMORe.
 *
 * This is the drop-in boundary. The reference has no FFI of its own: every model calls the
 * pronet core once per sample on host slices. Crossing a cgo / ctypes boundary per sample is impossible when the
 * rows live in HBM, so the seam moves up one level, to the model method set that every reference CLI uses
 *     New(); LoadEdgeList(file, undirected); Init(dim, ...); Train(...); SaveWeights(file)
 * (Go: cmd/line/main.go:54-67, cmd/bpr/main.go:42-55, cmd/deepwalk/main.go:44-57;
 *  C++: cli/line.cpp:71-76, cli/bpr.cpp:58-63, cli/warp.cpp:55-60, cli/hoprec.cpp:69-75, cli/deepwalk.cpp:72-77,
 *  cli/walklets.cpp:70-75). Each entry point below names the reference code it replaces.
 *
 * Conventions: plain pointers and sizes, no C++/torch types. Every call returns 0 on success or a negative
 * SMORE_E_* code; smore_last_error() returns the message for the calling thread. Handles own all device memory;
 * the caller keeps ownership of every host buffer it passes (the library copies). There is NO CPU fallback: every
 * compute call fails with SMORE_E_CUDA when no sm_100 device is usable.
 */
#ifndef SMORE_B200_H
#define SMORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smore_graph_s* smore_graph_t;
typedef struct smore_model_s* smore_model_t;

enum {
    SMORE_OK = 0,
    SMORE_E_INVALID = -1, /* bad argument */
    SMORE_E_CUDA = -2,    /* CUDA runtime error / no device */
    SMORE_E_IO = -3,      /* file could not be opened / parsed */
    SMORE_E_NOMEM = -4,
    SMORE_E_UNSUPPORTED = -5
};

/* Which of the two reference code bases a call reproduces (they differ: SURVEY.md §8c divergence table). */
enum { SMORE_SEM_CPP = 0, /* src/ + cli/  (proNet-core) */ SMORE_SEM_GO = 1 /* pkg/pronet + internal/models */ };

/* Negative-table weighting (C++: proNet::negative_method, src/proNet.cpp:486-509; BPR/WARP/HBPR ctors use no_degrees).
 * Go always uses DEGREES (pronet.go:242-249). */
enum { SMORE_NEG_DEGREES = 0, SMORE_NEG_IN_DEGREES = 1, SMORE_NEG_NO_DEGREES = 2 };

/* Execution mode. DETERMINISTIC = one warp consuming Philox stream `stream_base` sequentially: the reference at
 * -threads 1 fed the same draws. HOGWILD = grid-wide racing updates, warp w consumes stream `stream_base + w`. */
enum { SMORE_MODE_DETERMINISTIC = 0, SMORE_MODE_HOGWILD = 1 };

/* Stored element type of the embedding tables (the reference stores double). */
enum { SMORE_F32 = 0, SMORE_F64 = 1 };

/* Sampler selector for smore_graph_get_alias / smore_sample_debug. */
enum { SMORE_AT_VERTEX = 0, SMORE_AT_NEGATIVE = 1, SMORE_AT_CONTEXT = 2 /* C++ only */ };
enum { SMORE_SAMPLE_SOURCE = 0, SMORE_SAMPLE_NEGATIVE = 1, SMORE_SAMPLE_TARGET = 2, SMORE_SAMPLE_SOURCE_TARGET = 3 };

/* ---- library ---------------------------------------------------------------------------------------------------- */

/* Selects the CUDA device used by handles created afterwards by this thread/process (default: current device).
 * Fails unless the device is compute capability 10.x. */
int smore_init(int device_id);
const char* smore_last_error(void);
const char* smore_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t smore_kernel_launches(void);

/* ---- graph: replaces proNet::LoadEdgeList + BuildAliasMethod (src/proNet.cpp:115-236, :410-542) and
 *             ProNet.LoadEdgeList + buildGraph (pkg/pronet/pronet.go:112-249) ------------------------------------- */

/* From a host CSR in REFERENCE INSERTION ORDER (per-vertex adjacency in file order, duplicates kept, reverse entries
 * appended in place when undirected). row_off has V+1 entries. n_lines is the reference's MAX_line / MaxLine
 * (C++: E; Go: number of edge lines, not doubled); pass <=0 for E. The library builds the reference's alias tables
 * bit-exactly on the host (same Vose LIFO pairing order, same pow/normalisation expression) and uploads packed
 * integer-threshold tables. */
int smore_graph_create(int64_t V, int64_t E, const int64_t* row_off, const int32_t* col, const double* weight,
                       int64_t n_lines, int semantics, int negative_method, smore_graph_t* out);

/* Text ingest exactly as the reference does it: whitespace-separated `src dst weight`, ids in first-appearance order
 * (src before dst), malformed lines skipped (src/proNet.cpp:171-190; pronet.go:128-155). */
int smore_graph_load_edge_list(const char* path, int undirected, int semantics, int negative_method,
                               smore_graph_t* out);

/* The ingest alone, on the host (no device involved; multi-threaded, SMORE_HOST_THREADS): the CSR and the vertex names
 * smore_graph_load_edge_list would build. Two-call pattern: pass NULL arrays to get V, E, n_lines and names_bytes, then
 * arrays of V+1 / E / E entries and a names buffer (the names in id order, each terminated by '\n'). */
int smore_edge_list_to_csr(const char* path, int undirected, int64_t* V, int64_t* E, int64_t* n_lines, int64_t* row_off,
                           int32_t* col, double* weight, char* names, int64_t names_cap, int64_t* names_bytes);

/* HOP-Rec field metadata: proNet::LoadFieldMeta (src/proNet.cpp:330-408): `vertex field` lines, field ids in
 * first-appearance order; HOP-Rec treats field 0 as "user" (src/model/HBPR.cpp:98). */
int smore_graph_load_field(smore_graph_t g, const char* path);
int smore_graph_set_field(smore_graph_t g, const int32_t* field /* V entries */);

int smore_graph_info(smore_graph_t g, int64_t* V, int64_t* E, int64_t* n_lines);
/* Host copy of the CSR the graph was built from (ingest parity). Any pointer may be NULL. */
int smore_graph_get_csr(smore_graph_t g, int64_t* row_off, int32_t* col, double* weight);
/* NULL when the graph was created from a CSR (no names). */
const char* smore_graph_vertex_name(smore_graph_t g, int64_t vid);
/* Parity hook: the fp64 alias table as the reference holds it (AliasTable{alias, prob}, src/proNet.h:88-93). */
int smore_graph_get_alias(smore_graph_t g, int which, double* prob, int64_t* alias);
int smore_graph_get_field(smore_graph_t g, int32_t* field);
void smore_graph_destroy(smore_graph_t g);

/* Parity hook: run the DEVICE samplers on one warp over Philox stream (seed, stream), n draws, ids to `out`
 * (2n ids for SOURCE_TARGET). `arg` = source vertices for SMORE_SAMPLE_TARGET. Replaces SourceSample /
 * TargetSample(v) / NegativeSample (src/proNet.cpp:623-683; pronet.go:252-289). words_used may be NULL. */
int smore_sample_debug(smore_graph_t g, int which, uint64_t seed, uint64_t stream, int64_t n, const int64_t* arg,
                       int64_t* out, uint64_t* words_used);

/* Parity hook: device RandomWalk + SkipGrams (mode 0, window w0) / ScaleSkipGrams (mode 1, w0..w1) from `start`
 * on stream (seed, stream) (src/proNet.cpp:704-724, :769-809, :928-987; pronet.go:292-333). Writes the walk
 * (<= steps+1 ids) and up to `cap` (vertex, context) pairs; *n_pairs gets the true pair count. */
int smore_walk_debug(smore_graph_t g, uint64_t seed, uint64_t stream, int64_t start, int steps, int mode, int w0, int w1,
                     int64_t* walk, int64_t* walk_len, int64_t* pair_v, int64_t* pair_c, int64_t cap, int64_t* n_pairs);

/* ---- embedding store: replaces the models' vector<vector<double>> / [][]float64 tables
 *      (src/model/LINE.h:22-24, LINE.cpp:49-97; internal/models/line/line.go:44-70) ------------------------------- */

/* n_tables: 1 (LINE-1, BPR/WARP/HOP-Rec C++: one shared table) or 2 (vertex + context). Rows are dense [V x dim]. */
int smore_model_create(smore_graph_t g, int dim, int n_tables, int dtype, smore_model_t* out);
/* random != 0: (U - 0.5) / dim with U = k * 2^-32 from Philox stream (seed, 2^62 + table) -- the reference's
 * (rand()/RAND_MAX - 0.5)/dim (LINE.cpp:83) with a reproducible generator; random == 0: zeros (LINE.cpp:92). */
int smore_model_init(smore_model_t m, int table, int random, uint64_t seed);
int smore_model_set_rows(smore_model_t m, int table, int64_t first, int64_t n, const double* host);
int smore_model_get_rows(smore_model_t m, int table, int64_t first, int64_t n, double* host);
int smore_model_set_rows_f32(smore_model_t m, int table, int64_t first, int64_t n, const float* host);
int smore_model_get_rows_f32(smore_model_t m, int table, int64_t first, int64_t n, float* host);
/* Asynchronous variants for hosts that pipeline their own staging (fp32 models, `host` in pinned memory): uploads and
 * read-backs are queued on two internal copy streams, so the read-back of one Train() result overlaps the upload of the
 * next call's inputs (PCIe is full duplex). Within one model the two streams are ordered on the device PER TABLE: an upload
 * waits for the read-backs of the same table issued before it and a read-back for the earlier uploads of that table; copies
 * of different tables never wait on each other. The streams do not synchronise with the trainers: call
 * smore_model_wait_copies(m, uploads, readbacks) before training on rows being uploaded / before reading `host`. */
int smore_model_set_rows_f32_async(smore_model_t m, int table, int64_t first, int64_t n, const float* host);
int smore_model_get_rows_f32_async(smore_model_t m, int table, int64_t first, int64_t n, float* host);
int smore_model_wait_copies(smore_model_t m, int uploads, int readbacks);
/* Raw device pointer of a table (row-major [V x dim] of dtype) for zero-copy consumers on the same device. */
int smore_model_device_ptr(smore_model_t m, int table, void** ptr);
void smore_model_destroy(smore_model_t m);

/* ---- row sharding across the GPUs of one NVSwitch box: one process per GPU. New with this backend --
 * the reference is single-process shared memory (src/model/LINE.cpp:162 OpenMP workers on one table).
 *
 * Vertex v is owned by rank v & (world-1) (world in {1,2,4,8}); its local row is v >> log2(world). The rank that owns
 * the POSITIVE CONTEXT of a sample computes it: after smore_graph_set_shard the rank holds an alias table over the CSR
 * entries whose target it owns, weighted by P(source) * P(target | source) of the unsharded samplers, so the union over
 * ranks -- rank r running round(total * mass_r) samples, mass_r reported as source_mass_fraction -- reproduces the
 * global edge distribution exactly; negatives are drawn from the owned vertices only (stratified by the shard of the
 * positive, the PyTorch-BigGraph scheme). Context rows are therefore always local; the vertex row of a sample may
 * live on any rank and is reached with plain 128-bit loads/stores through CUDA-IPC peer mappings (NVLink). The CSR
 * stays replicated; smore_model_create allocates only the owned rows; set_rows / get_rows / init address LOCAL rows.
 * (Drawing negatives from the SOURCE's shard instead collapses the embedding: a vertex then never meets a negative
 * from another shard -- measured AUC 0.52 vs 0.92.) */
int smore_graph_set_shard(smore_graph_t g, int rank, int world);
int smore_graph_shard_info(smore_graph_t g, int* rank, int* world, int64_t* n_local, double* source_mass_fraction);
/* 64-byte cudaIpcMemHandle_t of this rank's shard of `table` (exchange with the host's own transport, e.g.
 * torch.distributed all_gather), then connect every peer: handles = world x 64 bytes, indexed by rank. */
int smore_model_ipc_handle(smore_model_t m, int table, void* handle64);
int smore_model_open_peers(smore_model_t m, int table, const void* handles);
/* Same-process variant (several shards on one device, used by tests): raw device pointers indexed by rank. */
int smore_model_set_peer_ptrs(smore_model_t m, int table, void* const* ptrs);

/* Optional read replica of a row-sharded table (table 0, the vertex table, for LINE): a full-size local copy indexed
 * by global vertex id. When enabled, smore_train_line reads (and locally updates) vertex rows in the replica and pushes
 * each row delta to the owner's shard with red.global.add over NVLink: no remote read on the sample's critical path and
 * half the NVLink bytes, at the price of staleness bounded by the refresh period. smore_model_refresh_replica re-pulls
 * every row from its owner; all ranks must be outside smore_train_* while any rank refreshes (host-side barrier). */
int smore_model_enable_replica(smore_model_t m, int table);
int smore_model_refresh_replica(smore_model_t m, int table);

/* Bulk-exchange mode of a row-sharded model (LINE): for graphs whose remote working set is too large for fine-grained
 * peer access (random 512-byte accesses to peer memory collapse once a few GB of peer rows are touched).
 * Training then proceeds in super-batches of `superbatch` samples per rank (<= 0: 2^20). Per super-batch a rank (1) files
 * the remote vertex rows its samples will need (the sampler is counter-based, so the sources are known before the
 * update runs) in per-owner request lists, (2) the lists, the requested rows and -- after the updates -- the modified
 * rows travel as contiguous buffers in three all-to-alls (NCCL send/recv groups over NVLink), (3) the owner adds
 * `returned - sent` to its row.
 * HOT vertices. A staged copy is updated in place by every sample of its rank during the super-batch (duplicates are
 * folded into one copy per rank), so what is stale is only the other ranks' view of the row, for one super-batch.
 * Measured on a 12 k-vertex graph in 4 shards (worst case: every vertex is drawn 11-44 times per super-batch): held-out
 * AUC 0.9207 / 0.9195 against 0.9218 unsharded. For extreme hubs the copies can still be avoided: a vertex expected to
 * be drawn as a source at least `hot_threshold` times per super-batch (all ranks together; P(source) of the unsharded
 * sampler x superbatch x world) keeps a single copy and is reached through the peer mappings exactly as in the
 * peer-access mode (smore_model_open_peers required); the hot set is small, so its peer footprint stays below the cliff.
 * hot_threshold < 0: no hot vertices (pure exchange, no peer mapping needed); 0: every vertex is hot (degenerates to the
 * peer-access mode). Recommended: 64.
 * Replaces nothing in the reference (its tables live in one address space); it is the "NCCL all-to-all carries the
 * remote-row batches" half of the sharded store, next to the peer-access half above. */
int smore_model_enable_exchange(smore_model_t m, int64_t superbatch, double hot_threshold);
/* NCCL bootstrap for the exchange mode with one process per GPU: rank 0 creates the 128-byte ncclUniqueId, the host
 * broadcasts it with its own transport, every rank calls init (after smore_init). libnccl.so.2 is dlopen()ed here. */
int smore_dist_nccl_unique_id(void* id128);
int smore_dist_nccl_init(const void* id128, int rank, int world);
int smore_dist_nccl_shutdown(void);
/* Debug hook: creates the exchange mode's SM-partitioned update stream with `reserve` SMs set aside (as
 * SMORE_EXCH_RESERVE_SMS would) and reports the device's SM count, the partition's, and on how many distinct SMs a probe
 * kernel launched on that stream really ran. First call in a process decides the partition. */
int smore_debug_sm_partition(int reserve, int* total_sms, int* partition_sms, int* sms_seen);
/* Counters of the last exchange-mode train call: super-batches run, remote vertex rows requested (after dedup); and the
 * size of the hot set. */
int smore_exchange_stats(smore_model_t m, uint64_t* superbatches, uint64_t* rows_requested, int64_t* hot_vertices);

/* ---- rotating shards: the multi-GPU mode for tables too large for fine-grained peer access ----------
 * New with this backend (the reference trains one shared table in one address space, src/model/LINE.cpp:162-191).
 * Both tables are row-sharded as above; the CONTEXT shard of a rank never moves. The VERTEX shard is cut into two halves
 * (local rows [0, sub_cap) and [sub_cap, n_local)), i.e. the table into 2*world sub-parts, which travel around the ring
 * of ranks: in episode e rank g trains the block  sources of sub-part q = (2g - e) mod 2*world  x  its own contexts
 * (negatives: its own contexts) entirely out of local HBM with the single-GPU kernel, while the sub-part it trained in
 * episode e-1 travels to rank g+1 as one contiguous copy-engine transfer over NVLink, hidden behind the update kernel.
 * After 2*world episodes (a cycle) every sub-part has met every context shard once and is home again. A vertex row
 * exists exactly once at any time (nothing stale, nothing lost); block (q, g) runs total_cycle * mass(q, g) samples per
 * cycle, so the union of the blocks is exactly the edge distribution of the unsharded samplers.
 *
 * Per rank:  smore_graph_create -> smore_graph_set_shard_rotating -> smore_model_create (2 tables) -> init / set_rows ->
 * smore_model_enable_rotation -> exchange slot handles with the NEXT rank (rot_ipc_handles / rot_open_next; same process:
 * rot_slot_ptrs / rot_set_next_ptrs) -> for e = 0, 1, ...:
 *        smore_rot_send_begin(m, e);  smore_train_line_episode(m, &p, e);  smore_rot_send_end(m, e);  <host barrier>
 * The barrier is the host's (torch.distributed / MPI / a Go channel): the library makes no cross-process call here.
 * get_rows / set_rows / init / checkpoints address LOCAL rows as for any sharded model and are valid whenever the model is
 * at home (episode % (2*world) == 0, after the barrier); smore_train_*_episode may be skipped in an episode (e.g. to
 * finish a partial cycle). Trainers: LINE order 2 and the Go BPR (models with separate vertex / context tables). */
int smore_graph_set_shard_rotating(smore_graph_t g, int rank, int world);
/* n_sub = 2*world; block_mass[q] / block_edges[q] (n_sub entries each, may be NULL): this rank's blocks by ring index q. */
int smore_graph_rotation_info(smore_graph_t g, int* n_sub, int64_t* sub_cap, double* block_mass, int64_t* block_edges);
int smore_model_enable_rotation(smore_model_t m);
/* 3 x 64-byte cudaIpcMemHandle_t of this rank's slot buffers; open the NEXT rank's ((rank + 1) % world) with open_next. */
int smore_model_rot_ipc_handles(smore_model_t m, void* handles192);
int smore_model_rot_open_next(smore_model_t m, const void* handles192);
/* Same-process variant (several shards on one device; tests): raw device pointers of the 3 slot buffers. */
int smore_model_rot_slot_ptrs(smore_model_t m, void** ptrs3);
int smore_model_rot_set_next_ptrs(smore_model_t m, void* const* ptrs3);
int smore_rot_send_begin(smore_model_t m, int64_t episode);
/* (smore_train_line_episode is declared with the trainers below) */

/* ---- device-side construction (smore_b200/csrc/device_graph.cu) --------------------------------
 * Parallel alias-table build on the GPU: the reference's AliasMethod (src/proNet.cpp:544-620, alias.go:10-90) is a
 * sequential two-stack sweep; this is the prefix-sum / binary-search formulation of the same sweep, O(n log n) work, no
 * sequential step. The table is a different valid alias table of the SAME distribution (another pairing order), so it is
 * used where no reference table exists to match: the block edge tables, shard-local negative tables and sub-part vertex
 * tables of the row-sharded store. thr[i] = ceil(prob_i * 2^32) (0xFFFFFFFF: the bucket never takes its alias). */
int smore_alias_build_device(const double* weights, int64_t n, uint32_t* thr, uint32_t* alias);
/* Synthetic power-law graph of BASELINE configs[4] shape, generated on the device straight into this rank's rotating-shard
 * tables: V vertices, E_lines undirected edge lines, line i = a pure function of (seed, i): endpoints with
 * P(rank r) ~ r^(-2/3) relabelled by a random bijection, integer weight U{1..5}, self loops dropped. No host CSR exists
 * (the reference could not even load such a graph: src/proNet.h:34, src/random.cpp:5): only the rotating-shard trainer
 * and the embedding store work on the returned handle. */
int smore_graph_create_synthetic_rotating(int64_t V, int64_t E_lines, uint64_t seed, int semantics, int rank, int world,
                                          smore_graph_t* out);
int smore_rot_send_end(smore_model_t m, int64_t episode);
int smore_rot_position(smore_model_t m, int64_t* episode, int* at_home, int* training_subpart);

/* Text writer: "<V> <dim>\n" then `name v0 v1 ...` per vertex in id order, VERTEX table only.
 * format 0 = C++ iostream default (%g, 6 significant digits; src/model/LINE.cpp:13-47),
 * format 1 = Go "%.6f" (internal/models/line/line.go:209-233). */
int smore_model_save_weights(smore_model_t m, int table, const char* path, int format);
/* Warm start: proNet::LoadPreTrain (src/proNet.cpp:238-286), the -load_v / -load_c flags of cli/deepwalk.cpp:61-62.
 * Text in the writer's own format; rows are matched by vertex NAME (the graph must come from an edge list), unknown
 * names and short lines are skipped, a dimension mismatch skips the whole file (0 rows, success) as the reference does.
 * On a row-sharded model each rank keeps the rows it owns. */
int smore_model_load_pretrain(smore_model_t m, int table, const char* path, int64_t* rows_loaded);
/* Binary snapshot of every table of this rank (raw rows in the stored element type + a header that pins V, dim, dtype,
 * rank/world, and -- format version 2 -- the training position, see smore_model_progress): save / resume for long runs -- the reference can import text embeddings but not resume. */
int smore_model_save_checkpoint(smore_model_t m, const char* path);
int smore_model_load_checkpoint(smore_model_t m, const char* path);
/* Where training stands (after the last train call, or as restored by load_checkpoint): the Philox key, the first sampler
 * stream no call has used yet, the length of the LR schedule and how much of it is done (units of params.total). To
 * resume, continue with the same seed, stream_base = next_stream, sched_total and sched_offset = sched_done; a version-1
 * checkpoint (tables only) restores 0 / 0 / 0 / 0: a warm start. Any pointer may be NULL. */
int smore_model_progress(smore_model_t m, uint64_t* seed, uint64_t* next_stream, uint64_t* sched_total, uint64_t* sched_done);
/* Live progress of the train call that is running on `m` -- the ONE entry point that may be called from another thread
 * while a smore_train_* call on the same handle blocks. Replaces the progress line the reference's workers
 * print every MONITOR samples (src/model/LINE.cpp:179-187, line.go:133-142: "Alpha: %.6f  Progress: %.3f %%"): warp 0 of
 * the training grid stores its view of the schedule -- units done (samples; walks for the walk models; this rank's share
 * on a row-sharded graph), the learning rate in force -- into host-mapped memory at every LR refresh, and this call only
 * reads that memory (no CUDA call, no lock). `total` is the length of the LR schedule (params.sched_total or .total),
 * `running` is 1 between the start of a train call and its return; after the call `done` is what the call really
 * finished. Before the first train call everything reads 0. Any pointer may be NULL. */
int smore_progress(smore_model_t m, uint64_t* done, uint64_t* total, double* alpha, int* running);
/* The writer's formatter on host rows (no device involved; multi-threaded): `<first_id + r> v0 v1 ...\n` per row, in the
 * number format of the chosen reference writer. Returns the byte count of the text (written to `out` when it fits in
 * `cap`; call with out = NULL to size the buffer) or a negative error. */
int64_t smore_format_rows(const double* rows, int64_t n, int dim, int64_t first_id, int format, char* out, int64_t cap);

/* ---- training: each call replaces one reference Train() body ------------------------------------------------------ */

typedef struct {
    int semantics;        /* SMORE_SEM_* (must match the graph's) */
    int mode;             /* SMORE_MODE_* */
    uint64_t seed;        /* Philox key */
    uint64_t stream_base; /* first sampler stream id */
    double alpha;         /* initial learning rate (-alpha) */
    uint64_t total;       /* C++: sample_times*1e6 (LINE.cpp:119) ; Go: sample_times*MaxLine (line.go:85) */
    int negative_samples; /* -negative_samples (LINE / skip-gram) */
    int order;            /* LINE: 1 or 2 */
    double lambda;        /* Go BPR L2 (bpr.go, -lambda); HPE / MF: -reg (cli/hpe.cpp:57, cli/mf.cpp:50) */
    int walk_times, walk_steps, window_min, window_max; /* DeepWalk: window_max = -window_size; Walklets: both */
    int max_warps;        /* HOGWILD: cap on concurrent warps (0 = fill the device) */
    int64_t max_walks;    /* DeepWalk/Walklets: stop after this many walks (<0: all walk_times*V) */
    /* Chunked training: this call is one piece of a longer LR schedule of sched_total units (same unit as `total`;
     * walk models: walks), sched_offset of which are already done. 0/0: the call is the whole schedule. */
    uint64_t sched_total, sched_offset;
    /* Skew-OPT (cli/skewopt.cpp:53-54): margin shift xi, scale omega, odd power eta */
    double xi, omega;
    int eta;
    /* Row-sharded LINE-2 (peer-access mode and rotating shards): how the K shard-local negatives of a sample are applied.
     * SMORE_PAIRING_AUTO / SMORE_PAIRING_COUPLED: the reference's pairing -- positive and negatives update the same vertex.
     * SMORE_PAIRING_SPLIT: split samples -- the K negatives update a second vertex drawn independently from the source
     * distribution, which makes the noise distribution of every vertex exactly the unsharded trainer's at the price of
     * one more row per update. Measured (profiles/r2*_ab_*): with atomic row updates both pass the 0.5 % AUC / recall@10
     * gate at 2 / 4 / 8 shards; on heavy-tailed graphs split keeps ~1 point more recall@10 when rotating-shard episodes
     * are long (many samples per vertex per episode). Ignored on unsharded graphs. */
    int neg_mode;
    /* node2vec (cmd/node2vec/main.go:21-22): return parameter p, in-out parameter q */
    double n2v_p, n2v_q;
    /* CPR (cmd/cpr/main.go:22-24): item regularisation and BPR margin (user regularisation = lambda); TPR (tpr.go:38,
     * cmd/tpr/main.go): weight of the text component of an item vector */
    double item_reg, margin, text_weight;
} smore_train_params;
#define SMORE_PAIRING_AUTO 0
#define SMORE_PAIRING_COUPLED 1
#define SMORE_PAIRING_SPLIT 2

/* Fills `p` with the reference CLI defaults (cmd/line/main.go:13-21, cli/line.cpp:56-64). */
void smore_train_params_default(smore_train_params* p);

/* LINE::Train (src/model/LINE.cpp:100-195) / LINE.Train (internal/models/line/line.go:73-150):
 * SourceSample -> TargetSample -> UpdatePair (1 + K context rows), LR decay every 10000 samples.
 * On a row-sharded model with the exchange mode enabled this call is COLLECTIVE: every rank must make it with the same
 * `total` (each runs its share in the same number of super-batches). */
int smore_train_line(smore_model_t m, const smore_train_params* p);
/* Exchange mode with every shard in the calling process (several shards on one device; tests and single-GPU
 * experiments): shards[r] must be rank r of n, all with the exchange mode enabled; plain device copies replace NCCL.
 * Shard r draws from streams stream_base + (r << 20) + warp. */
int smore_train_line_group(const smore_model_t* shards, int n, const smore_train_params* p);
/* Rotating shards (see above): one block of LINE::Train updates (src/model/LINE.cpp:162-191 restricted to the sources of
 * the resident vertex sub-part and the contexts of this rank). p->total = samples of ALL ranks in this episode (this rank
 * runs its block's share, total * 2*world * mass(q, rank)); p->sched_total / sched_offset in the same global unit;
 * p->stream_base must differ between ranks and episodes. */
int smore_train_line_episode(smore_model_t m, const smore_train_params* p, int64_t episode);
/* Same for the Go BPR (bpr.go:84-131, UpdateBPRPair optimizer.go:87-117): users live in the rotating vertex table, items in
 * the fixed context table; one block = users of the resident sub-part x items this rank owns, negatives among the rank's
 * own vertices. (The C++ BPR / WARP / HOP-Rec train ONE table in both roles and cannot be cut this way.) */
int smore_train_bpr_episode(smore_model_t m, const smore_train_params* p, int64_t episode);
/* BPR::Train (src/model/BPR.cpp:55-107, 5-negative UpdateBPRPair proNet.cpp:1406-1455) /
 * BPR.Train (internal/models/bpr/bpr.go:61-131, optimizer.go:87-117). */
int smore_train_bpr(smore_model_t m, const smore_train_params* p);
/* WARP::Train (src/model/WARP.cpp:55-107, UpdateWARPPair proNet.cpp:1353-1403). C++ only. */
int smore_train_warp(smore_model_t m, const smore_train_params* p);
/* HBPR::Train (src/model/HBPR.cpp:63-130, UpdateFBPRPair proNet.cpp:1458-1515). C++ only; needs field data. */
int smore_train_hoprec(smore_model_t m, const smore_train_params* p);
/* DeepWalk::Train (src/model/DeepWalk.cpp:98-155) / DeepWalk.Train (internal/models/deepwalk/deepwalk.go:61-141):
 * on-device RandomWalk + SkipGrams + UpdatePairs. */
int smore_train_deepwalk(smore_model_t m, const smore_train_params* p);
/* Node2Vec.Train (internal/models/node2vec/node2vec.go:82-260; Go tree only): DeepWalk.Train with the biased second-order
 * walk -- first step TargetSample, then per step weight * (1/p for the previous vertex | 1 for a neighbour of it | 1/q) summed
 * and scanned in adjacency order, one draw per step -- fixed full window, UpdatePairs. params.n2v_p / n2v_q. */
int smore_train_node2vec(smore_model_t m, const smore_train_params* p);
/* ---- CPR / TPR (Go tree only): models over TWO graphs and THREE tables ------------------------------------------------
 * The model is created on the FIRST graph (CPR: target domain, TPR: user-item graph; SMORE_SEM_GO) with two tables of V rows
 * -- table 0: users, table 1: items -- and the SECOND graph is attached as bare adjacency (CSR over its own vids, reference
 * insertion order) together with the third table of V_aux rows (CPR: source-domain item rows, only read; TPR: word rows,
 * trained). As in the reference a vertex of the first graph is looked up in the second BY ITS VID (cpr.go:148-169,
 * tpr.go:108): the caller must load the two edge lists so that shared entities get equal vids, exactly as it must for the
 * Go binaries. `rows` (V_aux x dim doubles) may be NULL: the table is then drawn (U - 0.5) / dim from `seed`.
 * Replaces CPR.LoadSourceDomain + Init (cpr.go:64-125) / TPR.LoadItemWordGraph + Init (tpr.go:49-99). */
int smore_model_attach_aux(smore_model_t m, int64_t V_aux, const int64_t* row_off, const int32_t* col, const double* rows,
                           uint64_t seed);
int smore_model_get_aux_rows(smore_model_t m, int64_t first, int64_t n, double* host);
/* CPR.Train (internal/models/cpr/cpr.go:175-282): SourceSample -> TargetSample -> transformUser (mean of the user row, the
 * rows of its target-domain neighbours and of its source-domain neighbours, :127-172) -> NegativeSample -> margin-gated BPR
 * step on the user row and the two target-domain item rows. params.total samples (update_times x 10^6), alpha, lambda =
 * user_reg, item_reg, margin. */
int smore_train_cpr(smore_model_t m, const smore_train_params* p);
/* TPR.Train (internal/models/tpr/tpr.go:124-262): BPR on text-enriched item vectors ((1 - text_weight) x item row +
 * text_weight x mean of the item's word rows, :101-121); updates the user row, both item rows and every word row of both
 * items. params.total samples (sample_times x MaxLine), alpha, lambda, text_weight. */
int smore_train_tpr(smore_model_t m, const smore_train_params* p);
/* HPE::Train (src/model/HPE.cpp:93-147): SourceSample -> TargetSample -> UpdateCommunity (src/proNet.cpp:3018-3054, the
 * context walks on for walk_steps steps; Opt_SigmoidRegSGD :1332-1351 with reg = params.lambda) -> UpdatePair with the
 * roles swapped. C++ only (the Go hpe model is LINE-2 with UpdatePair, hpe.go:97-107: use smore_train_line). */
int smore_train_hpe(smore_model_t m, const smore_train_params* p);
/* MF::Train (src/model/MF.cpp:50-98): SourceSample -> TargetSample -> UpdateFactorizedPair (src/proNet.cpp:2591-2614,
 * Opt_SGD :991-1012: linear prediction, labels +1/-1, reg = params.lambda) on table 0 in both roles. C++ only; create the
 * graph with SMORE_NEG_NO_DEGREES as the MF constructor does (MF.cpp:4-7). */
int smore_train_mf(smore_model_t m, const smore_train_params* p);
/* SPR::Train (src/model/SkewOPT.cpp): SourceSample -> TargetSample -> UpdateSBPRPair (src/proNet.cpp:1517-1566: 16 rounds,
 * every negative drawn inside; Opt_SBPRSGD :1070-1098 with params.xi / omega / eta) on table 0 in both roles. C++ only;
 * graph with SMORE_NEG_NO_DEGREES (SkewOPT.cpp:4-7). */
int smore_train_skewopt(smore_model_t m, const smore_train_params* p);
/* Walklets::Train (src/model/Walklets.cpp:6-64): RandomWalk + ScaleSkipGrams + UpdatePairs. C++ only. */
int smore_train_walklets(smore_model_t m, const smore_train_params* p);

/* Counters of the last train call on this model: samples (or walks) done, pair updates done, Philox words consumed
 * by stream `stream_base`, mean negatives scanned (WARP), device milliseconds of the kernels. Any may be NULL. */
int smore_train_stats(smore_model_t m, uint64_t* samples, uint64_t* pair_updates, uint64_t* words_stream0,
                      double* warp_mean_tries, double* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* SMORE_B200_H */
