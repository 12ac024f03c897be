/* jaeger_b200 -- C ABI of the B200-native `jaeger predict` hot path.
 *
 * Every entry point replaces one piece of the reference's Python/TensorFlow path
 * (paths relative to /root/reference/src/jaeger):
 *
 *   jg_fasta_scan/load   utils/fs.py:99-115, seqops/io.py:98-104 (pyfastx record iteration)    [host]
 *   jg_pack_bases        seqops/io.py:103-104 (.upper()) + encode.py:27-33 alphabet handling:
 *                        ASCII contig bytes -> 2-bit codes + validity bitmap in HBM
 *   jg_dust_mask         seqops/io.py:105-108 (pydustmasker sdust, window 64, threshold 20)
 *   jg_plan_windows      seqops/io.py:38-71 (_window_indices) + :112-145 (window loop,
 *                        is_last flags, short whole-contig windows)         [host integer code]
 *   jg_encode_windows    seqops/io.py:124-133 (G/C/A/T counts, gc_skew) +
 *                        seqops/encode.py:229-302 (process_string_inference) and
 *                        preprocess/v1/convert.py:75-99 (legacy process_string)
 *   jg_model_create /    nnlib/inference.py:311-339 (InferModel signature call) = the graph of
 *   jg_model_forward     nnlib/builder.py:844-894,982-1193 + nnlib/v2/layers.py (MaskedConv1D
 *                        1217-1280, MaskedBatchNorm 918-941, ResidualBlock 1882-1915,
 *                        MaskedGlobalMaxPooling 517-529, MaskedGlobalAvgPooling 460-480) +
 *                        nnlib/v2/nmd.py:43-77 + dense heads builder.py:589-596,705-713
 *   jg_aggregate_contigs postprocess/collect.py:293-403 (pred_to_dict per-contig reductions),
 *                        postprocess/helpers.py:175-235 (entropy, energy, sigmoid)
 *   jg_smooth_scores     postprocess/prophages.py:126-151 (softmax + width-4 box sum)
 *   jg_segment_scores    postprocess/prophages.py:554-595 (KernelCPD/PELT at pen 1..9)
 *   jg_viterbi_decode    postprocess/helpers.py:393-449 (--crf window decoding)
 *   jg_refine_contigs    postprocess/refinement.py:39-213 (--refine window labels + per-contig aggregation)
 *   jg_sw_scan / jg_sw_trace utils/termini.py:103-131 (terminal-repeat scans), postprocess/prophages.py:771-800 (att sites)
 *   jg_legacy_reliability postprocess/helpers.py:558-565 + collect.py:121-123 (legacy reliability_score)
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure, with a
 * message retrievable through jg_last_error().  Pointers named d_* are DEVICE pointers (the
 * Python host passes torch tensors' data_ptr()), h_* are host pointers.  All work is
 * stream-ordered on the context's stream; nothing synchronises unless stated.  There is no
 * CPU fallback: without a CUDA device jg_ctx_create fails.
 */
#ifndef JAEGER_B200_H_
#define JAEGER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jg_ctx jg_ctx;
typedef struct jg_model jg_model;

const char* jg_last_error(void);
int jg_version(void);

/* ---- context ------------------------------------------------------------------------- */
int jg_ctx_create(int device, jg_ctx** out);
/* Same, on a caller-owned stream (never destroyed by the library): the Python host passes a
 * torch.cuda.Stream so that torch's pinned-memory bookkeeping never sees a dead stream. */
int jg_ctx_create_on_stream(int device, void* stream, jg_ctx** out);
int jg_ctx_destroy(jg_ctx* ctx);
int jg_ctx_sync(jg_ctx* ctx);
/* raw cudaStream_t of the context (so the host can bracket work with its own events) */
void* jg_ctx_stream(jg_ctx* ctx);
/* number of CUDA kernels this library has launched on the context since creation */
int64_t jg_ctx_launch_count(jg_ctx* ctx);

/* ---- stage 0: FASTA ingest (host) ------------------------------------------------------------
 * Replaces the pyfastx passes of the reference (utils/fs.py:99-115, seqops/io.py:98-104): one
 * streaming pass over a plain or gzip-compressed file. jg_fasta_scan sizes the outputs; jg_fasta_load fills h_bases (n_bases bytes,
 * ideally pinned memory: it is the H2D source), h_offsets (n_records + 1) and h_names (record
 * names = header up to the first whitespace, NUL-terminated, back to back). */
int jg_fasta_scan(const char* path, int64_t* n_records, int64_t* n_bases, int64_t* name_bytes);
int jg_fasta_load(const char* path, uint8_t* h_bases, int64_t* h_offsets, char* h_names);

/* Chunked ingest for inputs that must not be held whole (BASELINE config 5, 100 Gbp; replaces the reference's record
 * iteration seqops/io.py:97-104).  jg_fasta_open: byte_begin / byte_end (-1 = to the end) give the reader exactly the records
 * whose '>' lies in [byte_begin, byte_end) of an UNCOMPRESSED file, so N ranks parse N disjoint byte slices (a gzip stream
 * can only be read from offset 0).  jg_fasta_next: the next run of whole records holding at most max_bases bases and
 * max_records records (at least one record); outputs as for jg_fasta_load.  *n_records == 0 marks the end.  Returns 3 with
 * *need_bases set when the next record alone does not fit cap_bases (it is kept for the next call). */
typedef struct jg_fasta_reader jg_fasta_reader;
int jg_fasta_open(const char* path, int64_t byte_begin, int64_t byte_end, jg_fasta_reader** out);
int jg_fasta_next(jg_fasta_reader* reader, int64_t max_bases, int64_t cap_bases, int64_t max_records,
                  int64_t cap_name_bytes, uint8_t* h_bases, int64_t* h_offsets, char* h_names,
                  int64_t* n_records, int64_t* n_bases, int64_t* name_bytes, int64_t* need_bases);
int jg_fasta_close(jg_fasta_reader* reader);

/* ---- stage 1: pack ---------------------------------------------------------------------
 * d_ascii: n bases (contigs concatenated, no separators).  d_codes: ceil(n/16) uint32, base i
 * in bits [2(i%16), 2(i%16)+2) with A=0 C=1 T=2 G=3 (complement = code ^ 2).  d_valid:
 * ceil(n/32) uint32, bit set when the (case-folded) base is A/C/G/T.  Both outputs must be
 * allocated with 2 extra zeroed words of slack. */
int jg_pack_bases(jg_ctx* ctx, const uint8_t* d_ascii, int64_t n, uint32_t* d_codes,
                  uint32_t* d_valid);

/* ---- stage 1b: low-complexity soft-mask (symmetric DUST, window 64) ------------------------------
 * Replaces pydustmasker.DustMasker(seq, window_size=64, score_threshold=threshold).mask()
 * (seqops/io.py:105-108).  The host cuts every contig into chunks; chunk i scans
 * [max(contig_begin, core_begin - 128), min(contig_end, core_end + 64)) and sets the bits of the
 * masked bases that fall inside [core_begin, core_end) in d_soft (same layout as d_valid, zeroed
 * by the caller).  All offsets are absolute base offsets into the packed arrays. */
int jg_dust_mask(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid,
                 const int64_t* d_core_begin, const int64_t* d_core_end,
                 const int64_t* d_contig_begin, const int64_t* d_contig_end, int64_t n_chunks,
                 int32_t threshold, uint32_t* d_soft);

/* ---- stage 2a: window plan (host) -------------------------------------------------------
 * For contig c of length len[c]: fixed stride windows range(0, len-(fsize-1), stride), or the
 * dynamic-stride placement when enabled; contigs with min_len <= len < fsize yield one
 * whole-contig window when `short_pass` is non-zero and are skipped otherwise; contigs longer
 * than max_len (when max_len > 0) are skipped.  Call with out pointers NULL to get the count.
 * Outputs (host, caller allocated, n_windows entries): contig index, window start,
 * window length in bases, ordinal within contig, is_last flag. */
int jg_plan_windows(const int64_t* h_len, int64_t n_contigs, int32_t fsize, int32_t stride,
                    int32_t dynamic_stride, double dynamic_stride_threshold, int32_t min_len,
                    int64_t max_len, int32_t short_pass, int64_t* n_windows,
                    int32_t* h_contig, int64_t* h_start, int32_t* h_nbases, int32_t* h_ordinal,
                    uint8_t* h_is_last);

/* ---- stage 2b: encode ------------------------------------------------------------------
 * d_win_base[w]: absolute base offset of window w in the packed arrays; d_win_nbases[w]: its
 * length (<= crop).  d_soft: optional soft-mask bitmap (same layout as d_valid) or NULL.
 * lut64: token of codon (b0*16 + b1*4 + b2) in the A0 C1 T2 G3 code; unknown codons -> 0.
 * case_sensitive != 0 makes soft-masked bases unknown for the tokens (reference masking=True
 * and the legacy encoder); counts always ignore soft-masked bases.
 * Outputs: d_tokens [n][6][pitch] uint8 (pitch >= lc, multiple of 4; zero padded), d_counts [n][4] int32 in G,C,A,T order
 * (meta_5..meta_8), d_skew100 [n] int16 = gc_skew*100 with Python round() semantics, bit 14
 * set when the value is a negative zero ("-0.000"). */
int jg_encode_windows(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid,
                      const uint32_t* d_soft, const int64_t* d_win_base,
                      const int32_t* d_win_nbases, int64_t n_windows, int32_t crop, int32_t lc,
                      int32_t pitch, const uint8_t* h_lut64, int32_t case_sensitive, uint8_t* d_tokens,
                      int32_t* d_counts, int16_t* d_skew100);

/* ---- stage 3: model ---------------------------------------------------------------------
 * The plan is a flat int32/float32 description compiled by the host from project.yaml; see
 * jaeger_b200/plan.py for the field layout.  Layer kinds: 1 = conv with fused epilogue, 2 =
 * MaxPooling1D(2) per frame, 3 = sum over frames + global max pool (legacy graph,
 * nnlib/v1/layers.py:65-69, 207, 413).  tok_offset: token t feeds one-hot channel t - tok_offset
 * (1 for the v2 encoder whose token 0 is the masked unknown codon, 0 for the legacy amino-acid
 * ids).  Weights are fp32 host arrays in TensorFlow layout ([k, Cin, Cout] conv kernels,
 * [in, out] dense kernels).  A conv layer's epilogue is: conv + bias -> norm 1 -> [+ shortcut] ->
 * activation 1 -> [NMD tap] -> [norm 2 -> activation 2] -> [masked global pool]; a norm is a
 * per-channel affine (MaskedBatchNorm folded, nnlib/v2/layers.py:918-941) or, with i[22] / i[23] set,
 * a MaskedDYT (layers.py:385-444): gamma * tanh(scale * x + shift) + beta with gamma / beta in
 * p[8..11].  i[28] != 0: the first norm is a MaskedLayerNormalization (layers.py:293-367) -- p[2] = gamma, p[3] = beta, p[1] the conv
 * bias (kept apart), i[29] = epsilon as float bits, statistics over the i[27] real channels.  i[30]: valid taps an output row needs
 * under mask propagation (layers.py:1245-1252): 0 / 1 "any", (k + 1) / 2 "majority", k "strict".  i[26] / i[27]: the layer's channel counts before the plan padded them to multiples of 64 (0 = not given).  A model
 * whose conv stack is at most 32 channels wide everywhere (BASELINE config 3) runs as ONE kernel that keeps each window in shared
 * memory through all layers (csrc/conv_resident.cuh); JG_RESIDENT=0 in the environment keeps the per-layer kernels. */
#define JG_LAYER_INT_FIELDS 32
#define JG_LAYER_PTR_FIELDS 16
typedef struct jg_layer_desc {
  int32_t i[JG_LAYER_INT_FIELDS];
  const float* p[JG_LAYER_PTR_FIELDS];
} jg_layer_desc;

typedef struct jg_head_desc {
  int32_t n_classes;        /* classifier units */
  int32_t feat_dim;         /* pooled feature width */
  int32_t pool_mode;        /* 1 max, 2 average */
  int32_t n_taps;           /* NMD taps (0 = no reliability head) */
  int32_t rel_hidden;       /* units of the reliability hidden dense (gelu) */
  int32_t mlp_hidden;       /* legacy head: two Dense(mlp_hidden, act) before the classifier (0 = none) */
  int32_t mlp_act;          /* activation code of those layers (3 = erf GELU) */
  int32_t reserved[1];      /* [0]: OOD signals appended to the NMD vector for the reliability head (reliability_model.mode
                             * "nmd_plus_signals", builder.py:644-657): n | id0 << 3 | id1 << 6 ...; ids 1 max_prob, 2 entropy,
                             * 3 energy, 4 margin, 5 nmd_norm (OODSignalLayer, nnlib/v2/layers.py:1632-1666); 0 = mode "nmd" */
  const float* cls_w;       /* [feat_dim][n_classes] */
  const float* cls_b;       /* [n_classes] */
  const float* rel_w1;      /* [sum tap widths + n signals][rel_hidden] */
  const float* rel_b1;      /* [rel_hidden] */
  const float* rel_w2;      /* [rel_hidden][1] */
  const float* rel_b2;      /* [1] */
  const float* mlp_w1;      /* [feat_dim][mlp_hidden] */
  const float* mlp_b1;
  const float* mlp_w2;      /* [mlp_hidden][mlp_hidden]; its output is the "embedding" */
  const float* mlp_b2;
} jg_head_desc;

int jg_model_create(jg_ctx* ctx, const jg_layer_desc* layers, int32_t n_layers,
                    const jg_head_desc* head, int32_t frames, int32_t tok_offset, jg_model** out);
int jg_model_destroy(jg_model* m);
/* Largest number of windows one jg_model_forward call may take for a given lc with the
 * workspace budget (bytes) -- the host chunks its window stream with it. */
int64_t jg_model_max_windows(jg_model* m, int32_t lc, int64_t workspace_bytes);
/* d_tokens [n][6][pitch] uint8; d_lpad [n] int32 = padded frame length of each window (lc for the
 * long pass; the batch maximum for the reference's padded short pass).  Outputs (device,
 * fp32): d_logits [n][n_classes]; optional d_rel [n], d_emb [n][feat_dim], d_nmd [n][sum taps]
 * (NULL to skip).  use_ref != 0 routes the convolutions through the CUDA-core check kernels
 * (tests only). */
int jg_model_forward(jg_ctx* ctx, jg_model* m, const uint8_t* d_tokens, const int32_t* d_lpad,
                     int64_t n_windows, int32_t lc, int32_t pitch, float* d_logits, float* d_rel,
                     float* d_emb, float* d_nmd, int32_t use_ref);
/* bytes of device workspace currently held by the model */
int64_t jg_model_workspace_bytes(jg_model* m);
/* useful FLOPs of one window at frame length lc (2*L_out*6*k*Cin*Cout summed over convs) */
double jg_model_flops_per_window(jg_model* m, int32_t lc);

/* Per-conv-launch timing with CUDA events on the context stream (bench.py's roofline leg).
 * jg_model_get_profile synchronises the stream and returns, per layer of the plan, the summed
 * kernel milliseconds, the number of launches and the number of windows they covered. */
int jg_model_set_profiling(jg_model* m, int32_t on);
/* "12 x jg::tc2::conv_tc2_kernel<2> + 4 x ..." : the conv kernel the last forward pass launched for every conv layer */
int jg_model_kernel_names(jg_model* m, char* buf, int32_t n);
int jg_model_get_profile(jg_model* m, int32_t n_layers, double* ms, int64_t* launches, double* windows);

/* ---- stage 4: per-contig aggregation ------------------------------------------------------
 * Windows of contig c are rows [d_offsets[c], d_offsets[c+1]) of d_logits [n_windows][n_cls].
 * Outputs per contig: mean/var as IEEE half bits (sequential fp32 accumulation then fp16
 * rounding, the arithmetic of np.mean / np.var on a float32 [T, C] array), consensus =
 * first-max argmax of the half means, per-class window counts of the per-window argmax,
 * entropy / energy means (half bits), number of windows with sigmoid(rel) > 0.5 (-1 when d_rel
 * is NULL; the host divides by the window count in double like np.mean), per-window argmax. */
int jg_aggregate_contigs(jg_ctx* ctx, const float* d_logits, const float* d_rel,
                         const int64_t* d_offsets, int64_t n_contigs, int64_t n_windows,
                         int32_t n_cls,
                         uint16_t* d_mean_h, uint16_t* d_var_h, int32_t* d_consensus,
                         int32_t* d_counts, uint16_t* d_entropy_h, uint16_t* d_energy_h,
                         int32_t* d_rel_pos, int32_t* d_frag_pred);

/* ---- stage 4b: prophage score smoothing + segmentation -------------------------------------
 * jg_smooth_scores: row softmax of logits (float64), then per class the width-`box` box SUM with
 * numpy.convolve(mode="same") alignment, per contig segment.  d_out [W][n_cls] float64. */
int jg_smooth_scores(jg_ctx* ctx, const float* d_logits, const int64_t* d_offsets,
                     int64_t n_contigs, int32_t n_cls, int32_t box, double* d_out);
/* jg_segment_scores: for one signal of n points, optimal partitioning with the L2 segment
 * cost, minimum segment length min_size and penalties pen = 1..n_pen (what
 * ruptures.KernelCPD(kernel="linear", min_size, jump=1).predict(pen=) minimises).
 * d_bkps [n_pen][n] int32: ascending segment ends for each penalty (last = n), d_nbkps [n_pen]. */
int jg_segment_scores(jg_ctx* ctx, const double* d_signal, int32_t n, int32_t min_size,
                      int32_t n_pen, int32_t* d_bkps, int32_t* d_nbkps);
/* The same search for every contig of a prophage-mode run in ONE launch pair (the reference loops over the
 * contigs in Python, postprocess/prophages.py:546-552): contig c owns points [d_offsets[c], d_offsets[c+1]) of the
 * concatenated signal (total_points in all).  d_bkps [n_pen][total_points]: contig c's list for penalty p starts at
 * d_bkps[p][d_offsets[c]]; d_nbkps [n_contigs][n_pen]. */
int jg_segment_scores_batched(jg_ctx* ctx, const double* d_signal, const int64_t* d_offsets, int32_t n_contigs,
                              int64_t total_points, int32_t min_size, int32_t n_pen, int32_t* d_bkps,
                              int32_t* d_nbkps);

/* Linear-chain CRF (Viterbi) decoding of every contig's window labels: replaces viterbi_decode
 * (postprocess/helpers.py:393-449) as called per contig from pred_to_dict (collect.py:343-346,
 * 367-375).  d_costs [n_cls][n_cls] float64 = build_transition_costs(...) (helpers.py:345-390).
 * Outputs: d_path [n_windows] decoded class per window, d_counts [n_contigs][n_cls]. n_cls <= 8. */
int jg_viterbi_decode(jg_ctx* ctx, const float* d_logits, const int64_t* d_offsets, int32_t n_contigs, int64_t n_windows,
                      int32_t n_cls, const double* d_costs, int32_t* d_path, int32_t* d_counts);

/* Post-hoc refinement (`jaeger predict --refine`): replaces add_score_features + refine + aggregate_contig
 * (postprocess/refinement.py:39-73, 97-137, 140-213) as driven by _build_refined_contig_df
 * (commands/predict.py:115-155).  Score columns are positional: 0 phage, 1 virus, 2 archaea, 3 bacteria,
 * 4 plasmid, 5 eukarya (n_cls >= 6).  d_tau [12] float64: per-class logit thresholds, then margin thresholds
 * (the calibration file's `taus`; -inf disables).  mode 0 gated / 1 weighted / 2 unweighted; merge_share 0.5
 * ("half") or 1.0 ("full").  Outputs: d_label [n_windows] (0..5 class, 6 unknown, 7 bacteria_or_plasmid,
 * 8 virus_any), d_margin [n_windows] float64, d_sums [n_contigs][6] float64, d_stats [n_contigs][2] =
 * {windows used, merged-label windows}, d_total_weight [n_contigs] float64. */
int jg_refine_contigs(jg_ctx* ctx, const float* d_logits, const int64_t* d_offsets, int64_t n_contigs, int64_t n_windows,
                      int32_t n_cls, const double* d_tau, int32_t merge_bp, int32_t merge_pv, int32_t mode, double merge_share,
                      uint8_t* d_label, double* d_margin, double* d_sums, int32_t* d_stats, double* d_total_weight);

/* Legacy `default` model reliability: replaces ood_predict_default(..., type "sklearn")
 * (postprocess/helpers.py:558-565; parameters loaded at commands/predict_legacy.py:99-109) per
 * window and np.mean per contig (collect.py:121-123).  coef / intercept are the logistic
 * regression's, cal_a / cal_b the prefit sigmoid calibrator's.  d_window_p0 [n_windows] float64,
 * d_contig_mean [n_contigs] float64 (may be NULL). */
int jg_legacy_reliability(jg_ctx* ctx, const float* d_embedding, int64_t n_windows, int32_t dim, const float* d_batch_mean,
                          const float* d_batch_std, const double* d_coef, double intercept, double cal_a, double cal_b,
                          const int64_t* d_offsets, int32_t n_contigs, double* d_window_p0, double* d_contig_mean);

/* ---- terminal-repeat scan -----------------------------------------------------
 * Replaces the two parasail.sw_trace_scan_16 calls per contig of scan_for_terminal_repeats
 * (utils/termini.py:103-131): matrix_create("ACGT", 2, -100), gap open 100 / extend 5.
 * A job aligns query = bases [q0, q0+nq) (nq = 0 means n) against reference = bases [r0, r0+n) (direct) or their
 * reverse complement (inverted), both taken from the packed contigs of jg_pack_bases.  Square jobs are the
 * terminal-repeat scans; rectangular ones the att-site scans of prophage_report (postprocess/prophages.py:771-800,
 * same parasail call and scores).
 * jg_sw_scan:  d_out[job] = {score, end_query, end_ref, length of the diagonal run ending there}.
 * jg_sw_trace: refills [0,end_i] x [0,end_j], one direction byte per cell at d_scratch + dirs_off
 *              ((end_i+1)*(end_j+1) bytes), and walks the traceback:
 *              d_out[job] = {alignment columns, gaps in query line, gaps in reference line, identities}
 *              (the counts get_alignment_summary takes from result.traceback, termini.py:43-46).  With
 *              ops_off >= 0 it also writes one byte per alignment column at d_scratch + ops_off, last column
 *              first (1 pair, 2 gap in the query line, 3 gap in the reference line; <= end_i + end_j + 2
 *              bytes): what result.traceback.query / .ref (the `front` / `rear`, `attL` / `attR` strings) need.
 * threads: CTA size, a multiple of 32 with threads * 16 >= max_rows (largest query length of the call, resp.
 * largest end_i + 1); max_cols = largest reference length, resp. largest end_j + 1. */
typedef struct jg_sw_job { int64_t q0, r0; int32_t n, inverted, nq, reserved; } jg_sw_job;
typedef struct jg_sw_trace_job { jg_sw_job job; int32_t end_i, end_j; int64_t dirs_off, ops_off; } jg_sw_trace_job;
int jg_sw_scan(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid, const jg_sw_job* d_jobs, int32_t n_jobs,
               int32_t threads, int32_t max_rows, int32_t max_cols, int32_t* d_out);
int jg_sw_trace(jg_ctx* ctx, const uint32_t* d_codes, const uint32_t* d_valid, const jg_sw_trace_job* d_jobs, int32_t n_jobs,
                int32_t threads, int32_t max_rows, int32_t max_cols, uint8_t* d_scratch, int32_t* d_out);

#ifdef __cplusplus
}
#endif
#endif /* JAEGER_B200_H_ */
