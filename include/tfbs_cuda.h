/*
 * tfbs_cuda.h — C ABI of libtfbs_cuda.so, the B200 (sm_100a) implementation of the
 * discover_tfbs feature-extraction hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference
 * (akshayparopkari/discover_tfbs) is pure Python, so the binding a maintainer adds is a ctypes
 * stub; discover_tfbs_b200/_lib.py is that stub.  Each entry point cites the
 * reference interface whose arithmetic it replaces (file:line in the reference repository).
 *
 * Conventions
 *   - every function returns 0 (TFBS_OK) or a negative tfbs_status; it never throws and never
 *     calls exit().  The message for the last failure on the calling thread is
 *     tfbs_last_error().  There is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with TFBS_ERR_NO_DEVICE / TFBS_ERR_CUDA.
 *   - the caller owns every host buffer (pointer + length); the library owns device memory
 *     behind opaque handles released by the matching *_destroy.
 *   - one host thread per tfbs_ctx; contexts on different devices may be driven concurrently
 *     (ctypes releases the GIL).  The only global state is the thread-local error string.
 *   - base coding: A=0 C=1 G=2 T=3, case-insensitive (acgt == ACGT, as Biopython's _pwm.c);
 *     every other byte (N, IUPAC ambiguity codes, gaps, separators) is INVALID.
 *     complement(b) = 3 - b  (reference: utils.py:97-105 reverse_complement).
 *   - a sequence SET (many FASTA records) is passed as one buffer with at least one invalid
 *     separator byte (e.g. '\n') between records; windows that touch a separator score NaN /
 *     never hit / give NA shapes, so no per-record bookkeeping is needed in the kernels.
 */
#ifndef TFBS_CUDA_H
#define TFBS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tfbs_status {
    TFBS_OK = 0,
    TFBS_ERR_INVALID = -1,    /* bad argument */
    TFBS_ERR_NO_DEVICE = -2,  /* no CUDA device / driver */
    TFBS_ERR_CUDA = -3,       /* CUDA runtime error (message has the CUDA string) */
    TFBS_ERR_NOMEM = -4,      /* device or host allocation failed */
    TFBS_ERR_OVERFLOW = -5    /* output capacity too small (n_hits still reports the need) */
} tfbs_status;

typedef struct tfbs_ctx tfbs_ctx;       /* one device + one stream + scratch */
typedef struct tfbs_seq tfbs_seq;       /* encoded sequence (2-bit words + invalid mask) in HBM */
typedef struct tfbs_pwms tfbs_pwms;     /* uploaded motif library (log-odds + derived LUTs) */
typedef struct tfbs_shapes tfbs_shapes; /* uploaded pentamer tables */

/* One thresholded hit.  pos is the 0-based start of the window on the + strand of the encoded
 * buffer; a reverse-strand hit stores ~pos (bitwise NOT, always negative) so that strand is
 * recoverable without a fifth field.  (Biopython's search() reports p - len(seq) for the
 * reverse strand; the Python wrapper converts.)  16 bytes. */
typedef struct tfbs_hit {
    int64_t pos;
    int32_t motif;
    float score;
} tfbs_hit;

#define TFBS_MAX_MOTIF_LEN 32     /* JASPAR-sized library: lengths 6..30 */
#define TFBS_SHAPE_PER_NUCLEOTIDE 0 /* MGW, ProT, EP : value = table[0][pentamer] */
#define TFBS_SHAPE_PER_STEP 1       /* Roll, HelT    : mean of Y2(P@k) and Y1(P@k+1) */

/* ---- library / context ---------------------------------------------------------------- */
const char *tfbs_last_error(void);
const char *tfbs_version(void);
int tfbs_device_count(int *n);
int tfbs_ctx_create(int device, tfbs_ctx **out);
int tfbs_ctx_destroy(tfbs_ctx *ctx);
int tfbs_ctx_sync(tfbs_ctx *ctx);
/* Device time (CUDA events on the context's stream) of the kernels launched by the most recent
 * compute call on this context, and how many kernels that call launched. */
int tfbs_last_kernel_ms(tfbs_ctx *ctx, float *ms);
int tfbs_last_kernel_launches(tfbs_ctx *ctx, int64_t *n);
/* Total kernels launched through this context since creation (bench.py's gpu_launches). */
int tfbs_total_kernel_launches(tfbs_ctx *ctx, int64_t *n);
/* How many filter candidates the most recent tfbs_scan_hits verified exactly (>= hits). */
int tfbs_last_scan_candidates(tfbs_ctx *ctx, int64_t *n);
/* Device time of the two kernels of the most recent tfbs_scan_hits: the filter scan and the exact
 * verify pass over its candidate list (bench.py reports a roofline for each). */
int tfbs_last_hits_ms(tfbs_ctx *ctx, float *scan_ms, float *verify_ms);
/* Of that scan time, the part spent in the direct-path kernel (0 when the path was not taken). */
int tfbs_last_direct_ms(tfbs_ctx *ctx, float *direct_ms);
/* Region timer: CUDA events recorded on the context's stream; stop synchronises and returns the
 * elapsed device time of everything enqueued in between (kernels and async copies). */
int tfbs_timer_start(tfbs_ctx *ctx);
int tfbs_timer_stop(tfbs_ctx *ctx, float *ms);
/* Pinned host memory for the end-to-end path (cudaHostAlloc / cudaFreeHost). */
int tfbs_host_alloc(void **ptr, int64_t bytes);
int tfbs_host_free(void *ptr);
/* Measured shared-memory load bandwidth of this GPU in GB/s (conflict-free LDS.32 streamed by 148 x 32
 * warps, best of 3): the denominator bench.py uses for the hits kernel's look-up-table roofline, next to
 * the nominal 148 SM x 128 B/clk x clock. */
int tfbs_measure_smem_gbs(tfbs_ctx *ctx, float *gbs);
/* Overwrite a >L2-sized scratch buffer so the next timed launch starts with a cold L2. */
int tfbs_flush_l2(tfbs_ctx *ctx);

/* ---- (1) sequence encoding -------------------------------------------------------------
 * Replaces the per-character Python string handling of utils.py:97-105 (reverse_complement),
 * utils.py:224-238 (calculate_gc_content) and the byte switch of Biopython's _pwm.c.
 * ASCII -> 2-bit words (16 bases / uint32, base i at bits 2*(i%16)) + invalid mask (32 bases /
 * uint32, bit i%32; padding bits past n are 1).  1.375 B/bp of HBM traffic. */
int tfbs_encode(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq **out);
/* Same, input already resident in HBM (device pointer). */
int tfbs_encode_device(tfbs_ctx *ctx, const void *ascii_dev, int64_t n, tfbs_seq **out);
/* Re-encode into an existing handle of the same length (no allocation; bench inner loop). */
int tfbs_encode_into(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq);
int tfbs_encode_device_into(tfbs_ctx *ctx, const void *ascii_dev, int64_t n, tfbs_seq *seq);
/* An empty handle for n bases (all invalid until tfbs_encode_into / tfbs_scan_hits_host fill it):
 * lets the end-to-end paths allocate once and re-use the device buffers for every shard or step. */
int tfbs_seq_create(tfbs_ctx *ctx, int64_t n, tfbs_seq **out);
int tfbs_seq_destroy(tfbs_seq *seq);
int tfbs_seq_length(const tfbs_seq *seq, int64_t *n);
/* Copy the encoding back: packed[ceil(n/16)], mask[ceil(n/32)] (parity tests). */
int tfbs_seq_download(tfbs_ctx *ctx, const tfbs_seq *seq, uint32_t *packed, uint32_t *mask);
/* counts[5] = #A #C #G #T #invalid over the whole buffer (popcounts on the packed planes). */
int tfbs_seq_base_counts(tfbs_ctx *ctx, const tfbs_seq *seq, int64_t counts[5]);
/* Per-record G+C counts for S records [start, start+len) of the encoded buffer
 * (utils.py:224-238: gc = (count('G') + count('C')) / len after upper()). */
int tfbs_gc_counts(tfbs_ctx *ctx, const tfbs_seq *seq, const int64_t *start, const int64_t *len,
                   int64_t S, int64_t *gc_out);
/* Device-side synthetic genome (position-addressable hash; identical to
 * discover_tfbs_b200/synth.py) written as ASCII into a library-owned HBM buffer. */
int tfbs_synth_ascii(tfbs_ctx *ctx, uint64_t seed, int64_t start, int64_t n, double gc,
                     int32_t with_n_runs, int32_t with_lowercase, void **ascii_dev_out);
int tfbs_device_free(tfbs_ctx *ctx, void *dev_ptr);
int tfbs_device_download(tfbs_ctx *ctx, const void *dev_ptr, void *host, int64_t bytes);

/* ---- (2) PWM log-odds scan, both strands -----------------------------------------------
 * Replaces Biopython PositionSpecificScoringMatrix.calculate / .search / .reverse_complement
 * (Bio/motifs/matrix.py, Bio/motifs/_pwm.c — the north-star oracle; the reference's own
 * candidate stage is exact-match BLASTn, utils.py:340-398, process_blast.sh:32).
 * logodds: double[M][Lmax][4], ACGT order, rows >= lens[m] ignored; -inf entries allowed.
 * 6 <= ... any 1 <= lens[m] <= TFBS_MAX_MOTIF_LEN. */
int tfbs_pwm_upload(tfbs_ctx *ctx, const double *logodds, const int32_t *lens, int32_t M,
                    int32_t Lmax, tfbs_pwms **out);
int tfbs_pwms_destroy(tfbs_pwms *p);
/* Dense mode: fwd[m][i], rev[m][i] = float32 score of motif m at start i on the + strand and of
 * its reverse-complement matrix at the same window (so rev[m][i] == fwd score of
 * revcomp(seq) at n-L-i); NaN where the window touches an invalid byte or runs past n.
 * Output layout [M][n] row-major.  Scores are accumulated exactly in per-motif scaled int32 fixed
 * point over double-precomputed k-mer partial sums and converted once to float, so they differ from
 * the double-accumulating oracle by at most one float rounding step of the score
 * (|error| <= 1e-4 for |score| < 1024, asserted in tests on adversarial sharp long motifs).
 * Host pointers may be NULL: results then stay in the context's device scratch (bench `value`). */
int tfbs_scan_dense(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_pwms *pwms, float *fwd_host,
                    float *rev_host);
/* Per-record best window (the PWM columns of the feature matrix; north_star item (3), joined like the
 * reference's per-site columns at utils.py:1134-1170): for S records [rec_start, rec_start + rec_len)
 * of the encoded buffer, best[s][m][0 fwd / 1 rev] = max over the record's windows of the score
 * calculate() gives (windows touching an invalid byte are skipped, numpy.nanmax semantics; NaN when
 * the record has no valid window or is shorter than the motif).  pos[s][m][strand] (may be NULL) =
 * offset of the first best window inside the record, -1 when none.  Every window is summed left to
 * right in double and cast to float, the maximum is taken over the floats: values and offsets are
 * bit-exact with max / first argmax of the oracle's calculate().  Only S*M*2 floats cross PCIe instead of the [M][n] dense arrays. */
int tfbs_scan_best(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_pwms *pwms, const int64_t *rec_start,
                   const int64_t *rec_len, int64_t S, float *best_host, int32_t *pos_host);
/* Hits mode: every (motif, start, strand) with (float)score >= thr[motif].  Candidates come from a packed
 * fixed-point filter over shared-memory k-mer tables with a rigorous upper bound (and, at sparse thresholds,
 * from an enumerated 12-mer hash for the motifs of at most 14 columns); a second kernel re-scores every
 * candidate in double — in the oracle's left-to-right order, or in column pairs when a rounding guard proves
 * that both orders round to the same float — so hit sets and scores are bit-exact with the oracle.  Order
 * of `out` is unspecified unless `sorted` is non-zero (then ascending (motif, pos, strand)).  out_host may
 * be NULL (hits stay on device).  If more than `cap` hits exist the call returns TFBS_ERR_OVERFLOW with
 * *n_hits = the number found. */
int tfbs_scan_hits(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_pwms *pwms, const float *thr,
                   tfbs_hit *out_host, int64_t cap, int32_t sorted, int64_t *n_hits);

/* End-to-end form for a HOST-resident buffer (ideally pinned, tfbs_host_alloc): `seq` is a handle
 * of length n (e.g. from a previous tfbs_encode) whose contents are overwritten.  The buffer is cut
 * into `chunks` ranges; H2D copy, encode + scan and the D2H copy of each range's hits overlap on
 * three streams.  Same results as tfbs_encode_into + tfbs_scan_hits. */
int tfbs_scan_hits_host(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq,
                        const tfbs_pwms *pwms, const float *thr, tfbs_hit *out_host, int64_t cap,
                        int32_t chunks, int32_t sorted, int64_t *n_hits);
/* Shard form of tfbs_scan_hits_host for chunk + halo sharding across GPUs: the buffer
 * is one shard [shard_start, shard_start + n) of a longer sequence, passed WITH its right halo of
 * (Lmax - 1) bases; only windows that start in its first `owned` bases are reported (the halo's
 * windows belong to the next shard) and shard_start is added to every position — both on the device,
 * so the per-shard outputs concatenate into the whole-genome result without a host-side filter.
 * With `sorted` the shard's hits are ordered (motif, pos, strand) on the device before the copy. */
int tfbs_scan_hits_host_shard(tfbs_ctx *ctx, const uint8_t *ascii_host, int64_t n, tfbs_seq *seq,
                              const tfbs_pwms *pwms, const float *thr, tfbs_hit *out_host, int64_t cap,
                              int32_t chunks, int32_t sorted, int64_t owned, int64_t shard_start,
                              int64_t *n_hits);
/* PCI bus id of a device ("0000:1b:00.0"): the host layer reads the device's NUMA-local CPU list from
 * /sys/bus/pci/devices/<id>/local_cpulist and pins the thread that allocates and fills the pinned
 * buffers of that GPU to it, so that 8 GPUs' copies do not all cross to one socket's memory. */
int tfbs_device_pci_bus_id(int device, char *buf, int32_t len);
/* Page-lock caller-owned host memory in place (cudaHostRegister) so the copies of the end-to-end
 * paths run at full PCIe rate without a staging copy; undo with tfbs_host_unregister. */
int tfbs_host_register(void *ptr, int64_t bytes);
int tfbs_host_unregister(void *ptr);
/* The blocks the last hits scan of this library used: narrow (64 motifs, 8-bit filter fields) or wide
 * (32 motifs, 16-bit) is chosen per block from the thresholds by a cost model (expected candidates
 * against shared-memory loads saved).  out[4b..4b+3] = {A, B, narrow, motifs}; `cap` = blocks `out`
 * holds (out may be NULL to query *n_blocks).  Diagnostics for bench.py's shared-memory roofline. */
int tfbs_pwms_hits_blocks(const tfbs_pwms *pwms, int32_t *out, int32_t cap, int32_t *n_blocks);

/* The DIRECT path of the last hits scan: at sparse thresholds the motifs of at most 14 columns are not
 * filtered through the shared-memory tables; every L-mer that reaches the threshold is enumerated on the
 * host into a 12-mer hash behind a 10-mer bitmap and the scan looks positions up in it.  Chosen per motif
 * and per threshold vector.  n_motifs = how many motifs took it (the blocks of tfbs_pwms_hits_blocks cover
 * the rest), max_len = the longest of them, keys = 12-mers in the hash; all 0 when the path was not taken
 * (dense thresholds, very short motifs, fewer than 32 eligible motifs, or TFBS_HITS_DIRECT=0). */
int tfbs_pwms_hits_direct(const tfbs_pwms *pwms, int32_t *n_motifs, int32_t *max_len, int64_t *keys);

/* ---- (3) DNA-shape pentamer lookup -----------------------------------------------------
 * Replaces DNAshapeR getShape()/getDNAShape() (get_DNA_shapes.R:23-25,
 * create_bkg_seqs.py:365-366) and the text parsing + window slicing the reference does on its
 * output (build_features.py:193-198,215-241; utils.py:1061-1084 _get_DNA_shape).
 * tables: float[T][2][1024]; pentamer index = sum code[k]*4^(4-k); kinds[t] is
 * TFBS_SHAPE_PER_NUCLEOTIDE (uses tables[t][0]) or TFBS_SHAPE_PER_STEP (Y1 = tables[t][0],
 * Y2 = tables[t][1]). */
int tfbs_shapes_upload(tfbs_ctx *ctx, const float *tables, const int32_t *kinds, int32_t T,
                       tfbs_shapes **out);
int tfbs_shapes_destroy(tfbs_shapes *s);
/* Whole-buffer tracks: out[t][i], NaN = NA.  Records are delimited by invalid separator bytes;
 * rec_start/rec_len (S records, may be NULL/0 = the buffer is one record) give the record
 * boundaries needed for the single-pentamer step values next to record ends.
 * out_host may be NULL (tracks stay in device scratch). */
int tfbs_shape_tracks(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_shapes *shapes,
                      const int64_t *rec_start, const int64_t *rec_len, int64_t S,
                      float *out_host);
/* Per-site feature matrix out[s][t][w], w < W, written exactly as the reference slices it:
 *   mode 0 (utils.py:1061-1084): site = (record, start, end, strand);
 *           '+': indices start..end, '-': indices end..start (descending), i.e. L+1 values with
 *           L = end-start; an index whose token would be 'NA' or the trailing '' of the
 *           flattened file (index == n_tokens when trailing_empty) yields 0.0; an index past the
 *           tokens yields NaN (missing column); '-' with start == 0 yields an all-NaN row.
 *   mode 1 (build_features.py:215-241): background records of length len: per-nucleotide
 *           tracks give indices 2..len-3, per-step tracks 1..len-3; NA -> 0.0.
 * rec_start/rec_len locate the record (chromosome) of every site inside the encoded buffer. */
int tfbs_shape_sites(tfbs_ctx *ctx, const tfbs_seq *seq, const tfbs_shapes *shapes,
                     const int64_t *rec_start, const int64_t *rec_len, const int64_t *start,
                     const int64_t *end, const int8_t *strand, int64_t S, int32_t W, int32_t mode,
                     int32_t trailing_empty, float *out_host);

/* Legacy input path: the caller already holds DNAshapeR output as per-record token lists (the
 * reference's {shape: {chrom: [str, ...]}} dict, build_features.py:193-198) parsed to floats
 * ('NA' and '' -> 0.0, as utils.py:1076-1077 does).  Gathers the reference's Python slices
 * ('+': [start : end+1], '-': [end : start-1 : -1]) of list rec_off[s] .. +rec_ntok[s] into
 * out[s][w]; NaN where the slice is shorter than W. */
int tfbs_track_sites(tfbs_ctx *ctx, const float *values_host, int64_t n_values,
                     const int64_t *rec_off, const int64_t *rec_ntok, const int64_t *start,
                     const int64_t *end, const int8_t *strand, int64_t S, int32_t W, float *out_host);

/* ---- (4) similarity features ("next" row) ------------------------
 * All-pairs sums over a true set of U sequences for S candidates, replacing the Python loops of
 * utils.py:1043-1058 over jaccard_similarity (utils.py:1243-1255) and pac (utils.py:892-949):
 * out[s] = { sum_u jaccard(true_u, cand_s), sum_u PAS(true_u, cand_s), sum_u PPS(true_u, cand_s) }
 * (the caller divides by U for the reference's np.mean).  Per sequence the host passes the sorted
 * unique k-mer codes (2 bits per base, first base most significant) with multiplicities, the
 * total k-mer count, and the sorted unique MurmurHash3_x86_32(seed 39) values of the canonical
 * k-mers (utils.py:1259-1269); *_woff / *_hoff are CSR offsets.  One k per call. */
int tfbs_similarity_sums(tfbs_ctx *ctx, const uint32_t *c_words, const uint16_t *c_counts,
                         const int32_t *c_woff, const int32_t *c_nk, const uint32_t *c_hash,
                         const int32_t *c_hoff, int64_t S, const uint32_t *t_words,
                         const uint16_t *t_counts, const int32_t *t_woff, const int32_t *t_nk,
                         const uint32_t *t_hash, const int32_t *t_hoff, int64_t U, double *out_host);

/* ---- (5) scaler + SVC decision function ("next" row) -----------------
 * Replaces the inference arithmetic of predict.py:286-319 for a FITTED model: x1 = x * mm_scale + mm_min
 * (MinMaxScaler), x2 = (x1 - st_mean) / st_scale (StandardScaler; cross_validate_model.py:182-189), then
 * sklearn SVC.decision_function: kernel 0 (linear): sv = coef_ [F], out = coef_ . x2 + intercept;
 * kernel 1 (rbf): out = sum_i dual[i] * exp(-gamma * |sv[i] - x2|^2) + intercept, sv = [n_sv][F].
 * X: double[S][F] raw features; out: double[S].  All arithmetic in double. */
int tfbs_svc_decision(tfbs_ctx *ctx, const double *X_host, int64_t S, int32_t F, const double *mm_scale,
                      const double *mm_min, const double *st_mean, const double *st_scale, int32_t kernel,
                      const double *sv, const double *dual, int32_t n_sv, double gamma, double intercept,
                      double *out_host);

/* ---- test hooks (host-only, no device needed) -------------------------------------------
 * The encode kernel converts 4 ASCII bytes per 32-bit word with integer bit tricks; this runs
 * the same inline function on the host so the CPU test-suite can check it exhaustively. */
int tfbs_debug_swar4(uint32_t four_ascii_bytes, uint32_t *codes8, uint32_t *valid4);
/* The dense kernel's fixed-point k-mer tables of ONE motif (logodds double[L][4], ACGT): k-mer width
 * K, groups G, scale exponent e (entries = rint(group sum * 2^e)), the bound below which a window sum
 * means -inf (INT32_MIN when the motif has no -inf entry), and optionally the entries themselves as
 * int32[G][4^K][2] (fwd, rev; `cap` = int32 slots `entries` holds).  A window score is
 * (float)(sum over groups) * 2^-e: lets the CPU suite replay the kernel's arithmetic on adversarial
 * matrices against the double-accumulating oracle (Biopython _pwm.c semantics). */
int tfbs_debug_dense_tables(const double *logodds, int32_t L, int32_t *K, int32_t *G, int32_t *scale_exp,
                            int32_t *neg_thr, int32_t *entries, int64_t cap);
/* Column grouping of the hits kernel for a motif block of maximum length L: A groups of 4 columns
 * + B groups of 3 (the kernel is instantiated for every pair this can return for L = 1..32). */
int tfbs_debug_hits_group_plan(int32_t L, int32_t *A, int32_t *B);
/* Blocks the hits kernel cuts a library of M motif lengths (any order) into, longest first, when every
 * block of at most narrow_max_groups column groups is taken in narrow form:
 * out[4b..4b+3] = {A, B, narrow, motifs in block b}; `out` holds 4 * M int32, *n_blocks is set.
 * A narrow block carries 64 motifs (8-bit filter fields), a wide block 32 (16-bit fields). */
int tfbs_debug_hits_block_plan(const int32_t *lens, int32_t M, int32_t narrow_max_groups, int32_t *out,
                               int32_t *n_blocks);

/* The candidate FILTER of the hits kernel run on the host with the kernel's own tables and 32-bit
 * field arithmetic (no CUDA): form 0 = wide blocks, 1 = narrow, 2 = narrow + sticky, 3 = the plan a
 * scan would use (cost model) with every motif on the tables, 4 = the table side of the plan a scan
 * really uses (motifs on the direct path are left out and get no candidates here).  codes = n bases as 2-bit codes, one per byte; cand = uint8[n][M][2],
 * 1 where window start p of motif m passes on strand s.  Optional outputs: plan = int32[blocks][4]
 * {A, B, form, motifs} (room for M blocks), expected = double[blocks] candidates per window start the
 * table builder predicts for a narrow / sticky block, n_blocks.  Lets the CPU suite check that the
 * filter never loses a hit (Biopython search semantics) in any block form and that
 * the cost model's prediction matches what the filter passes. */
int tfbs_debug_hits_filter(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                           const float *thr, int32_t form, const uint8_t *codes, int64_t n, uint8_t *cand,
                           int32_t *plan, double *expected, int32_t *n_blocks);

/* The DIRECT path of the hits scan planned and looked up on the host (no CUDA): codes = n bases as 2-bit
 * codes, one per byte; cand = uint8[n][M][2], 1 where the 12-mer at p is an enumerated key of (motif,
 * strand) — which, for a window of valid bases inside the sequence, is exactly "the oracle reports a hit".
 * info = {motifs on the direct path, their longest length, keys, bitmap bits set}. */
int tfbs_debug_direct_lookup(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                             const float *thr, const uint8_t *codes, int64_t n, uint8_t *cand, int64_t info[4]);
/* Which motifs the direct path takes for a threshold vector: is_direct[M] (0 / 1).  The choice is made per
 * motif — one whose own 12-mer keys are too many (loose or -inf threshold, very short motif) stays on the
 * table filter and the rest of the library still takes the path. */
int tfbs_debug_direct_motifs(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                             const float *thr, uint8_t *is_direct);

/* The COMPLETE plan a hits scan makes for a threshold vector (direct-path choice, block cut, block forms,
 * every filter table), built on the host by the code the scan itself runs and reduced to checksums:
 * out = {blocks, table words, FNV-1a of the filter tables + initial field values, FNV-1a of the block
 * descriptors, motifs on the direct path, direct-path keys}; plan (optional, room for M blocks) =
 * int32[blocks][4] {A, B, form, motifs}.  Pins the plan across refactorings and host thread counts. */
int tfbs_debug_hits_plan_checksum(const double *logodds, const int32_t *lens, int32_t M, int32_t Lmax,
                                  const float *thr, uint64_t out[6], int32_t *plan);

#ifdef __cplusplus
}
#endif
#endif /* TFBS_CUDA_H */
