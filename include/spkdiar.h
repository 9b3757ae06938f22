/*
 * spkdiar.h - C-ABI of libspkdiar.so, the B200 (sm_100a) implementation of the
 * statistical core of the Aalto speaker-diarization scripts.
 *
 * The reference (JianCao92/speaker-diarization) has NO plugin / FFI interface:
 * its hot path is three Python-2 scripts whose numeric kernels are in-script
 * functions calling numpy / scipy.  This header is therefore the boundary WE
 * define; each entry point names the reference
 * function(s) whose work it replaces.  The host side that binds it is
 * speaker-diarization_b200/_abi.py (ctypes); INTEGRATION.md shows the stub a
 * maintainer of the reference scripts would add.
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success or a negative
 *     SPKDIAR_E_* code, with text from spkdiar_last_error();
 *   - the caller owns every host buffer; the library owns device memory behind
 *     opaque handles; handles belong to the context that made them;
 *   - one context per (host thread, device); calls on one context are
 *     serialised by the caller; every call is synchronous on return;
 *   - frame positions are frame indices into the uploaded feature matrix,
 *     half-open ranges [a, b); distances are IEEE fp64;
 *   - there is no CPU fallback: without a usable sm_100-class device
 *     spkdiar_create() fails.
 */
#ifndef SPKDIAR_H
#define SPKDIAR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPKDIAR_ABI_VERSION 1

/* error codes */
#define SPKDIAR_OK            0
#define SPKDIAR_E_CUDA       -1   /* a CUDA runtime call failed                     */
#define SPKDIAR_E_ARG        -2   /* bad argument (null pointer, range, dimension)  */
#define SPKDIAR_E_NOMEM      -3   /* host or device allocation failed               */
#define SPKDIAR_E_CAPACITY   -4   /* caller-provided output buffer too small        */
#define SPKDIAR_E_UNSUPPORTED -5  /* feature dimension / option not built           */
#define SPKDIAR_E_NODEVICE   -6   /* no CUDA device / not an sm_100-class device    */

/* distance measures: -d GLR|BIC|KL2 (spk-change-detection.py:428-435) */
#define SPKDIAR_GLR 0
#define SPKDIAR_BIC 1
#define SPKDIAR_KL2 2

/* the feature dimension the kernels are compiled for (fconfig.cfg:78-83).  Feature files of 1..SPKDIAR_DIM
 * dimensions are accepted (the reference's loader takes the dimension from the file header,
 * spk-change-detection.py:37-41): the library keeps such frames zero-padded to SPKDIAR_DIM columns and treats
 * the padding as an identity block, so every distance is that of the file's own dimensions; one dimension per
 * device at a time (handles of another dimension are refused while any is alive). */
#define SPKDIAR_DIM 39
/* doubles per sufficient-statistics record: 780 packed second moments,
 * 39 first moments, 1 frame count */
#define SPKDIAR_RECORD 820

typedef struct spkdiar_ctx  spkdiar_ctx;
typedef struct spkdiar_feat spkdiar_feat;
typedef struct spkdiar_clus spkdiar_clus;

/* ---- context --------------------------------------------------------------- */

int  spkdiar_abi_version(void);
/* device: CUDA ordinal.  stream: a cudaStream_t the caller already owns (e.g.
 * torch's current stream) passed as void*, or NULL to let the context create
 * its own non-blocking stream. */
int  spkdiar_create(int device, void* stream, spkdiar_ctx** out);
void spkdiar_destroy(spkdiar_ctx* ctx);
/* last error text of ctx (or of the failed spkdiar_create when ctx is NULL) */
const char* spkdiar_last_error(const spkdiar_ctx* ctx);
/* number of kernels the library launched on this context since creation */
int64_t spkdiar_launch_count(const spkdiar_ctx* ctx);
/* number of SMs of the context's device */
int  spkdiar_sm_count(const spkdiar_ctx* ctx);

/* Kernel-class timers (CUDA events on the context's stream).  enable != 0
 * switches timing on and clears the accumulators.  spkdiar_profile_read fills
 * ms[k] / launches[k] for k < SPKDIAR_NPROF. */
#define SPKDIAR_PROF_STATS   0   /* frame-statistics prefix kernels (K1)            */
#define SPKDIAR_PROF_SCORE   1   /* batched window / pair scoring (K2, K6)          */
#define SPKDIAR_PROF_GW      2   /* persistent growing-window driver (K3)           */
#define SPKDIAR_PROF_MERGE   3   /* persistent agglomerative merge loop (K7)        */
#define SPKDIAR_PROF_H2D     4   /* host->device feature copies                     */
#define SPKDIAR_NPROF        5
int  spkdiar_profile_enable(spkdiar_ctx* ctx, int enable);
int  spkdiar_profile_read(const spkdiar_ctx* ctx, double* ms, int64_t* launches);

/* ---- features + frame statistics (K1) ---------------------------------------
 * Replaces load_features' in-memory result (spk-change-detection.py:31-43) as
 * the operand of every later np.cov / np.mean: the (n, dim) float32 frame-major
 * matrix is copied to HBM and fp64 prefix sums of x and x x^T are built so that
 * the covariance of any window is an O(dim^2) difference. */
int  spkdiar_features_upload(spkdiar_ctx* ctx, const float* frames, int64_t n,
                             int32_t dim, spkdiar_feat** out);
/* same, for a matrix that already lives in device memory (not copied, must
 * outlive the handle) */
/* The same WITHOUT the window statistics (the 6,560 B-per-frame prefix records of K1): frames and shift only.
 * What spk-clustering.py needs when it runs on its own (get_spk_features, spk-clustering.py:46-52: the frames
 * of every turn, nothing per window): spkdiar_cluster_create / _run then accumulate the cluster records straight
 * from the frames (156 B read per frame).  Any entry point that scores windows (spkdiar_gw_run,
 * spkdiar_score_windows, spkdiar_score_sets, spkdiar_stats_window, spkdiar_cluster_batch) builds the
 * statistics on first use, after which the handle behaves like one from spkdiar_features_upload. */
int  spkdiar_features_upload_frames(spkdiar_ctx* ctx, const float* frames, int64_t n,
                                    int32_t dim, spkdiar_feat** out);
int  spkdiar_features_adopt(spkdiar_ctx* ctx, const float* dev_frames, int64_t n,
                            int32_t dim, spkdiar_feat** out);
/* A BATCH of recordings in one handle (BASELINE config 4: spk-diarization2.py over a corpus,
 * lines 122-128 once per media file).  Recording r is copied to the packed frame rows
 * [base_out[r], base_out[r] + n[r]) (base_out[r] is a multiple of 128) and its statistics
 * restart there, so every result over a packed recording is bit-identical to the result over
 * the same recording uploaded alone; frame positions passed to later calls on this handle
 * (chains, segments, windows) are PACKED positions.  One spkdiar_gw_run over the chains of all
 * recordings and one spkdiar_cluster_batch then process the whole batch in two launches. */
int  spkdiar_features_upload_batch(spkdiar_ctx* ctx, const float* const* frames, const int64_t* n,
                                   int32_t nrec, int32_t dim, spkdiar_feat** out, int64_t* base_out);
/* (re)build the prefix statistics; upload/adopt already call it once */
int  spkdiar_stats_build(spkdiar_feat* f);
int  spkdiar_features_free(spkdiar_feat* f);
int64_t spkdiar_features_frames(const spkdiar_feat* f);
/* test hook: sufficient statistics of frames [a, b) in NATURAL order:
 * out[0..38] = sum (x - shift), out[39 + i*(i+1)/2 + j] = sum (x-shift)_i (x-shift)_j
 * (j <= i), shift[0..38] = the per-file shift the library subtracted */
int  spkdiar_stats_window(spkdiar_feat* f, int64_t a, int64_t b, double* out819,
                          double* shift39);

/* ---- batched window scoring (K2) --------------------------------------------
 * Replaces bic / glr / kl2 (spk-change-detection.py:72-133) for ncand
 * candidates at once.  Candidate k splits [a[k], b[k]) at m[k]: left = [a, m),
 * right = [m, b), pooled = [a, b).
 *   out_d[k]        the distance (BIC uses lambda);
 *   out_terms       NULL, or 3 doubles per candidate:
 *                     BIC: ln|S_left|, ln|S_right|, ln|S_pooled|
 *                     GLR: ln|S_left|, ln|S_right|, ln|(N1/N) S1 + (N2/N) S2|
 *                     KL2: first trace term, second trace term, 0
 *                   (the host needs the terms to replay the reference's
 *                    mutable-default BIC memo, spk-change-detection.py:72,84-90)
 */
int  spkdiar_score_windows(spkdiar_feat* f, const int64_t* a, const int64_t* m,
                           const int64_t* b, int64_t ncand, int metric,
                           double lambda, double* out_d, double* out_terms);

/* Distance between two SETS of frame ranges (clusters): set 1 = ranges
 * [a1[i], b1[i]) i < n1, set 2 likewise; pooled = concatenation.  Replaces one
 * dist(arr1, arr2) call of spk-clustering.py:136-175 (in-order clustering) and
 * of merge_rec (spk-change-detection.py:136-177).  npairs problems are scored
 * at once: problem p uses ranges off1[p]..off1[p+1] of (a1, b1) and
 * off2[p]..off2[p+1] of (a2, b2). */
int  spkdiar_score_sets(spkdiar_feat* f, int64_t npairs,
                        const int64_t* off1, const int64_t* a1, const int64_t* b1,
                        const int64_t* off2, const int64_t* a2, const int64_t* b2,
                        int metric, double lambda,
                        double* out_d, double* out_terms);

/* ---- growing-window search (K3) ---------------------------------------------
 * Replaces dist_gw (spk-change-detection.py:180-288): the whole sequential
 * search (coarse scan, threshold test, fine tune, reset-at-change, window
 * growth) runs on the device, one independent chain per recipe line. */
typedef struct {
    double rate;        /* float(-f)                                   CD:499 */
    double winsize;     /* floor(-w * rate)      frames                CD:526 */
    double winstep;     /* floor(-st * rate)     frames                CD:527 */
    double deltaws;     /* floor(rate * -dws)    frames                CD:508 */
    double threshold;   /* -t                                          CD:532 */
    double lambda;      /* -l (BIC only)                               CD:455 */
    int32_t metric;     /* SPKDIAR_GLR | _BIC | _KL2                          */
    int32_t max_groups; /* 0 = automatic; else cap on concurrently running
                           chains (1 = the whole GPU works on one chain)      */
} spkdiar_gw_params;

/* one record per window the reference's outer loop visits (CD:201) */
typedef struct {
    double  start;      /* window start, chain-relative, fp64 as the reference */
    double  end;        /* window end                                          */
    double  maxi;       /* best coarse offset i                                */
    double  maxd;       /* best coarse distance; -2^63 when no candidate won   */
    double  maxi_fine;  /* after the fine tune (positive windows only)         */
    double  maxd_fine;
    int32_t positive;   /* 1: a change was written at start + maxi_fine        */
    int32_t chain;      /* index of the recipe line / chain                    */
    int32_t ncand;      /* coarse candidates evaluated; when none of them won
                           (all NaN / inf / no candidate) encoded as -(n + 1)  */
    int32_t ninf;       /* candidates (coarse + fine) whose distance was +-inf */
    int32_t seq;        /* position of the window within its chain             */
    int32_t pad;
} spkdiar_gw_window;

/* chains: chain c covers frames [seg_a[c], seg_b[c]) of f.  Window records are
 * returned chain by chain, in window order; win_first[c]..win_first[c+1]
 * (nchain + 1 entries) index chain c's records.  A too small win_cap returns
 * SPKDIAR_E_CAPACITY and stores the needed count in win_first[0]. */
int  spkdiar_gw_run(spkdiar_feat* f, const spkdiar_gw_params* params,
                    const int64_t* seg_a, const int64_t* seg_b, int32_t nchain,
                    spkdiar_gw_window* win, int64_t win_cap, int64_t* win_first);

/* nrun searches over the SAME chains side by side (BASELINE config 2: the BIC, GLR and
 * KL2 passes over one recipe): each search is one dependent chain of waves that is far
 * from filling the GPU, so they run concurrently on disjoint subsets of the SMs.
 * params[r], win[r], win_cap[r], win_first[r] belong to search r; results are exactly
 * those of nrun separate spkdiar_gw_run calls. */
#define SPKDIAR_MAX_RUNS 4
int  spkdiar_gw_run_multi(spkdiar_feat* f, int32_t nrun, const spkdiar_gw_params* params,
                          const int64_t* seg_a, const int64_t* seg_b, int32_t nchain,
                          spkdiar_gw_window* const* win, const int64_t* win_cap,
                          int64_t* const* win_first);

/* The same searches as an asynchronous object.  begin() launches them (searches that run split into
 * sub-chains share one stream and one set of SMs: the first starts at once, the others when they are
 * waited for); wait(run) blocks until search `run` is complete and fills its records as spkdiar_gw_run
 * does (SPKDIAR_E_CAPACITY: wait again with a larger buffer); where(run) tells on which stream and how
 * many SMs that search ran, so that work depending on it (spkdiar_ctx_exec + the clustering of its
 * turns) can be queued there while the other searches are still running; end() releases the object
 * (and waits for searches nobody collected).  One such object per context at a time. */
typedef struct spkdiar_gwm spkdiar_gwm;
int  spkdiar_gw_multi_begin(spkdiar_feat* f, int32_t nrun, const spkdiar_gw_params* params,
                            const int64_t* seg_a, const int64_t* seg_b, int32_t nchain, spkdiar_gwm** out);
int  spkdiar_gw_multi_wait(spkdiar_gwm* h, int32_t run, spkdiar_gw_window* win, int64_t win_cap,
                           int64_t* win_first);
int  spkdiar_gw_multi_where(const spkdiar_gwm* h, int32_t run, void** stream, int32_t* sms);
int  spkdiar_gw_multi_end(spkdiar_gwm* h);
/* Later calls on ctx run on `stream` (a cudaStream_t) with at most `sms` SMs; (NULL, 0) restores the
 * context's own stream and the whole device. */
int  spkdiar_ctx_exec(spkdiar_ctx* ctx, void* stream, int32_t sms);

/* Host-only self-test of the sub-chain split / stitch / continuation logic behind spkdiar_gw_run (no
 * device needed): a toy search whose state after a change depends on `start` alone is cut into
 * sub-chains, stitched round by round by the real code and compared with the toy's sequential search.
 * p_sync: probability that searches from different starts agree on a change.  0 = passed. */
int  spkdiar_selftest_stitch(uint64_t seed, double p_sync, int64_t nframes, int32_t nsub_target);

/* ---- agglomerative clustering (K5-K7) ---------------------------------------
 * Replaces spk_cluster_hi of spk-clustering.py:178-260 (variant 1) and of
 * spk-clustering2.py:173-229 (variant 2).  Initial clusters are the frame
 * ranges [seg_a[k], seg_b[k]).  The device keeps per-cluster sufficient
 * statistics and the pair matrix resident and runs the merge loop without a
 * host round trip.  metric = SPKDIAR_BIC | _GLR | _KL2; for KL2 (spk-clustering.py:124-133) the
 * engine caches diag(S), diag(S^-1) and the running float32 sum of every cluster (a merged
 * cluster's sum continues a's over b's frames turn by turn, as np.mean of the concatenation
 * does) - single GPU, spkdiar_cluster_run only. */
typedef struct {
    int32_t a;          /* surviving cluster, index in the compacted list      */
    int32_t b;          /* removed cluster (a < b), same indexing              */
    double  d;          /* the minimum that triggered the merge                */
} spkdiar_merge;

/* stats[0..3] = max_dist, min_dist, max_det_dist, min_det_dist as the scripts
 * track them (variant 1: over every finite distance ever computed,
 * spk-clustering.py:196-200,233-237,210-213; variant 2: matrix max / min at
 * convergence, spk-clustering2.py:220-221), starting from 0 / 2^63-1. */
int  spkdiar_cluster_create(spkdiar_feat* f, const int64_t* seg_a,
                            const int64_t* seg_b, int64_t nseg, int metric,
                            double lambda, spkdiar_clus** out);
int  spkdiar_cluster_run(spkdiar_clus* c, double threshold, int32_t max_spk,
                         int32_t variant, spkdiar_merge* out, int64_t cap,
                         int64_t* nmerges, double* stats4);
/* sharded run for one very long recording (BASELINE config 5), variant 1: the pair
 * matrix is dealt out over nranks devices, pair (r, c) is scored and kept by rank
 * (r + c) % nranks (every row of the matrix is spread evenly, so the rescoring of a
 * merged row is balanced); cluster statistics are replicated, so every rank applies
 * the same merges.  All ranks must have created the handle from the same segments.
 * exchange() is called once per merge with this rank's best candidate (16 bytes:
 * fp64 distance, int64 flat index r * nseg + c) and must fill all[nranks] with
 * every rank's candidate (an all-gather; the host supplies it - torch.distributed,
 * MPI, threads - so that the library does not link a communication library); it is
 * called once more at the end with (max, min) of the distances for stats4.  The
 * global minimum is taken with ndarray.argmin's order (NaN first, value, flat index),
 * which makes the merge sequence identical to spkdiar_cluster_run's for any nranks.
 * nranks == 1 needs no exchange function. */
typedef int (*spkdiar_exchange_fn)(void* user, const void* mine16, void* all);
int  spkdiar_cluster_run_sharded(spkdiar_clus* c, double threshold,
                                 int32_t max_spk, int32_t rank, int32_t nranks,
                                 spkdiar_exchange_fn exchange, void* user,
                                 spkdiar_merge* out, int64_t cap,
                                 int64_t* nmerges, double* stats4);
/* The same run with the exchange done by NCCL on the context's stream (libnccl.so.2 is
 * loaded at run time, e.g. the copy torch ships): merges are queued in batches of
 * argmin kernel -> ncclAllGather of the 16-byte candidates -> deciding + rescoring kernel,
 * the host only looks at a stop flag between batches.  Rank 0 creates the id with
 * spkdiar_nccl_unique_id() and the caller distributes its 128 bytes to all ranks. */
int  spkdiar_nccl_unique_id(void* out128);
int  spkdiar_cluster_run_sharded_nccl(spkdiar_clus* c, double threshold, int32_t max_spk,
                                      int32_t rank, int32_t nranks, const void* unique_id128,
                                      spkdiar_merge* out, int64_t cap, int64_t* nmerges,
                                      double* stats4);
/* The same run as ONE persistent kernel per rank whose ranks exchange their candidates
 * through peer memory (NVLink): every rank owns a mailbox of 3 x nranks 32-byte slots
 * (candidates of even merges, of odd merges, and the final statistics round, which therefore never shares a
 * slot with the first candidate of the NEXT run on the same mailboxes);
 * per merge, CTA 0 of a rank stores its candidate into its slot of EVERY rank's mailbox
 * (payload, then a sequence number with st.release.sys) and all CTAs poll their own
 * device's mailbox (ld.acquire.sys) - no host round trip, no collective library call.
 * mailboxes[r] is rank r's mailbox as mapped on this device (spkdiar_mailbox_open of the
 * handle rank r made with spkdiar_mailbox_create; mailboxes[rank] is the local pointer).
 * seq_base: every rank passes the same value, larger by at least nseg + 2 than in the
 * previous run on the same mailboxes.  The ranks' kernels must run at the same time, on
 * different devices; a rank that waits longer than ~1e8 polls gives up with an error. */
int  spkdiar_mailbox_create(spkdiar_ctx* ctx, int32_t nranks, void** local_ptr, void* ipc_handle64);
int  spkdiar_mailbox_open(spkdiar_ctx* ctx, const void* ipc_handle64, void** peer_ptr);
int  spkdiar_mailbox_close(spkdiar_ctx* ctx, void* peer_ptr);
int  spkdiar_mailbox_free(spkdiar_ctx* ctx, void* local_ptr);
int  spkdiar_cluster_run_sharded_p2p(spkdiar_clus* c, double threshold, int32_t max_spk,
                                     int32_t rank, int32_t nranks, void* const* mailboxes,
                                     uint64_t seq_base, spkdiar_merge* out, int64_t cap,
                                     int64_t* nmerges, double* stats4);
/* Many SMALL clustering problems in one launch (one CTA per problem, no grid barrier):
 * problem p clusters the segments first[p] .. first[p + 1] of (seg_a, seg_b) exactly as
 * spkdiar_cluster_create + spkdiar_cluster_run would (same merges, distances, statistics,
 * bit for bit).  out has first[nprob] slots, the merges of problem p start at out[first[p]];
 * nmerges[p] of them are valid; stats4 (may be NULL) holds 4 doubles per problem.  Problems
 * larger than 4096 segments are refused (SPKDIAR_E_UNSUPPORTED): use the resident engine. */
int  spkdiar_cluster_batch(spkdiar_feat* f, int32_t nprob, const int64_t* first,
                           const int64_t* seg_a, const int64_t* seg_b, int metric, double lambda,
                           double threshold, int32_t max_spk, int32_t variant,
                           spkdiar_merge* out, int64_t* nmerges, double* stats4);
/* In-order clustering, spk_cluster_in of spk-clustering.py:136-175 / spk-clustering2.py:135-170, with the loop over
 * the recipe lines ON THE DEVICE (one launch per recording instead of one scoring call per line): line l is the frame
 * range [seg_a[l], seg_b[l]); it is scored against every speaker found so far (BIC or GLR of the speaker's frames and
 * the segment's), joins the nearest one when that distance is <= threshold, else becomes a new speaker.  The speakers
 * that exist before the first line (an earlier wav of the same recipe; nspk0 may be 0) are given as range sets:
 * speaker s = ranges off0[s] .. off0[s + 1] of (a0, b0).  Out: dist[dist_first[l] .. dist_first[l + 1]) = the
 * distances of line l to the speakers 0, 1, ... that existed then (none for the very first speaker), best[l] = the
 * speaker joined or -1 for a new one.  The host replays the script's bookkeeping (statistics, -tt lines) from the
 * distances.  A too small dist_cap returns SPKDIAR_E_CAPACITY (nlines * (nspk0 + nlines) always suffices). */
int  spkdiar_cluster_inorder(spkdiar_feat* f, int32_t nspk0, const int64_t* off0, const int64_t* a0,
                             const int64_t* b0, int64_t nlines, const int64_t* seg_a, const int64_t* seg_b,
                             int metric, double lambda, double threshold,
                             double* dist, int64_t dist_cap, int64_t* dist_first, int32_t* best);
/* Merge mode of the change detector, merge_rec of spk-change-detection.py:136-177 (driver loop 375-394), with the
 * loop over the recipe lines of ONE wav on the device: step k scores the previous segment - line 0, or the run of
 * lines merged so far, as ONE range from the first line's start to the last merged line's end - against line k + 1
 * (BIC or GLR) and merges when the distance is below the threshold.  seg_a / seg_b: the clamped frame bounds
 * int(start * rate), int(end * rate) of every line.  BIC with use_memo != 0 reproduces the reference's memo of the
 * FIRST left term: *memo_c1 in (NaN = not set yet) and out.  Out per step: the three ln|S| terms
 * (left, right, pooled / mixed), the distance and whether the lines were merged. */
int  spkdiar_merge_chain(spkdiar_feat* f, int64_t nlines, const int64_t* seg_a, const int64_t* seg_b,
                         int metric, double lambda, double threshold, int32_t use_memo, double* memo_c1,
                         double* terms, double* dist, int32_t* merged);
int  spkdiar_cluster_free(spkdiar_clus* c);
/* test hook: copy the current pair matrix (nseg x nseg, row-major, entries of
 * dead rows / columns undefined) and the alive flags to the host */
int  spkdiar_cluster_matrix(spkdiar_clus* c, double* out, uint8_t* alive);
/* measurement hook: SM-clock cycles CTA 0 spent in the five phases of the last persistent merge loop on this
 * handle (spkdiar_cluster_run, _run_sharded_p2p), summed over the merges: out6[0] cached-minimum scan,
 * [1] first grid barrier, [2] pick + exchange with the other ranks + stop test, [3] rescoring of the merged row,
 * [4] second grid barrier; out6[5] = merges.  bench.py reports them per merge for the sharded configuration. */
int  spkdiar_cluster_counters(const spkdiar_clus* c, uint64_t* out6);
/* test hook: the NEXT spkdiar_cluster_run on this handle also copies row `a` of the pair matrix as merge m
 * rewrote it (nseg doubles, ORIGINAL indices; entries of dead columns undefined) to host_rows[m * nseg ...],
 * for the first cap_rows merges - what a host needs to replay the agglomeration and check that every merge
 * was ndarray.argmin of the live matrix (spk-clustering.py:203-205).  cap_rows = 0 switches it off. */
int  spkdiar_cluster_rowlog(spkdiar_clus* c, double* host_rows, int64_t cap_rows);

/* ---- host replay of spk-diarization2.py's two calls, per recording (no device work) -----------
 * Replaces the text side of `spk-change-detection.py ... -m gw` + `spk-clustering.py -m hi` as
 * spk-diarization2.py:122-128 chains them: recipe parsing (spk-change-detection.py:11-28), one chain per run
 * of equal lna (370-374), a recipe line per detected change (252-256, 286-288 through write_recipe_line,
 * 46-69: LNA renaming, Python-2 str(float)), the clustering stage reading that text back into one cluster per
 * line (spk-clustering.py:263-283), merges applied to the speaker lists and the turns written smallest first
 * (216-220, 243-260).  The window records / merges come from spkdiar_gw_run / spkdiar_cluster_batch.
 * Anything this unit does not reproduce to the byte (non-ASCII text, exponents or signs in a time field,
 * several wavs in one recipe) is answered with SPKDIAR_E_UNSUPPORTED and text from spkdiar_replay_error():
 * the caller then runs the general (Python) replay of that recording.  Handles are independent of contexts
 * and of each other (any thread). */
typedef struct spkdiar_replay spkdiar_replay;
/* parses the recipe text (lines separated by '\n').  *out is set whenever memory could be had - also together
 * with SPKDIAR_E_UNSUPPORTED (read the message, then free it). */
int  spkdiar_replay_create(double rate, const char* recipe_text, int64_t len, spkdiar_replay** out);
int  spkdiar_replay_free(spkdiar_replay* r);
const char* spkdiar_replay_error(const spkdiar_replay* r);
/* out6 = parsed lines, chains, 1 if all lines name one wav, turns written by _segment, speakers after
 * _cluster, windows replayed */
int  spkdiar_replay_info(const spkdiar_replay* r, int64_t* out6);
/* the chains spkdiar_gw_run is to search: [base + a, base + b) per chain, frames clamped to nframes */
int  spkdiar_replay_chains(const spkdiar_replay* r, int64_t nframes, int64_t base, int64_t* seg_a,
                           int64_t* seg_b, int64_t cap);
/* window records of this recording's chains: chain k = win[win_first[k] .. win_first[k + 1]) */
int  spkdiar_replay_segment(spkdiar_replay* r, const spkdiar_gw_window* win, const int64_t* win_first);
/* the initial clusters of the clustering stage, one per line of the segmentation recipe */
int  spkdiar_replay_turns(const spkdiar_replay* r, int64_t nframes, int64_t base, int64_t* seg_a,
                          int64_t* seg_b, int64_t cap);
/* the merge sequence (compacted indices, as spkdiar_cluster_run returns them) -> clustered recipe */
int  spkdiar_replay_cluster(spkdiar_replay* r, const spkdiar_merge* merges, int64_t nmerges);
/* which = 0: segmentation recipe (after _segment), 1: clustered recipe (after _cluster); owned by r */
const char* spkdiar_replay_text(const spkdiar_replay* r, int32_t which, int64_t* len);

#ifdef __cplusplus
}
#endif
#endif /* SPKDIAR_H */
