/*
 * genofix_b200 -- C ABI of the B200-native Mendelian-rank / genotype-correction hot path.
 *
 * Drop-in boundary for GenoFix (andreaskranis/genofix).  The reference is pure Python and has no
 * FFI; the entry points below are what a ctypes binding inside the reference's own modules would
 * call (INTEGRATION.md shows the stubs).  Each entry point cites the reference code it replaces
 * (paths relative to the reference's src/ directory).
 *
 * Conventions
 *   - every function returns 0 on success, a negative GFX_E_* code on failure; the message is
 *     available from gfx_last_error() (thread local).  No exceptions cross the ABI.
 *   - pointers named *_dev are CUDA device pointers on the current device; *_host are host
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - genotypes on the device are 2-bit packed: animal-major rows, 16 codes per uint32 word,
 *     SNP j of a row lives in word j/16 at bits 2*(j%16); code 0/1/2 = genotype, 3 = missing
 *     (the reference's 9).  `wpr` = words per row (>= ceil(n_snps/16), a multiple of 4 so that rows are
 *     16-byte aligned); padding codes must be 3.
 *   - rank matrices are IEEE binary16 bit patterns (uint16_t), row stride `ld_raw` elements.
 *   - handles are opaque and owned by the caller until the matching *_destroy.
 */
#ifndef GENOFIX_B200_H
#define GENOFIX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFX_OK 0
#define GFX_E_INVALID (-1)   /* bad argument                                                   */
#define GFX_E_CUDA (-2)      /* CUDA runtime error (message has the cudaError string)          */
#define GFX_E_PEDIGREE (-3)  /* pedigree outside the reference's domain (e.g. one known parent, */
                             /* pedigree_dag.py:126-128 KeyError)                    */
#define GFX_E_TOO_WIDE (-4)  /* a blanket needs more simultaneous free variables than GFX_WMAX  */
#define GFX_E_NOMEM (-5)

typedef struct gfx_schedule gfx_schedule;

const char* gfx_last_error(void);
int gfx_version(void);
/* Number of kernels launched by this library in the calling process since load (bench `gpu_launches`). */
int64_t gfx_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Pedigree schedule: replaces PedigreeDAG indexes + per-animal get_parents_depth / get_subset /
 * balance_nodes / generateModel (pedigree/pedigree_dag.py:124-168,188-231; bayesnet/model.py:173-230)
 * and correctMatrix's bookkeeping (correct_genotypes3.py:556-565,684-686,851-853).
 *
 * All arrays are HOST arrays indexed by pedigree animal 0..n_ped-1:
 *   ped_id      original id (canonical pair order is ascending (sire id, dam id))
 *   sire, dam   pedigree index of the parent or -1
 *   generation  min(pure sire-line length, pure dam-line length)   (pedigree_dag.py:197-200)
 *   row         genotype-matrix row of the animal or -1 if not genotyped
 *   kid_off/kid children CSR in the reference's iteration order (set(successors), see
 *               correct_genotypes3.py:199-200); n_ped+1 offsets
 * back = blanket depth for rank and singles; pairs use back+1 (correct_genotypes3.py:691).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_ped;
  int32_t n_rows;
  int32_t back;
  const int64_t* ped_id;
  const int32_t* sire;
  const int32_t* dam;
  const int32_t* generation;
  const int32_t* row;
  const int32_t* kid_off;
  const int32_t* kid;
} gfx_pedigree_desc;

int gfx_schedule_create(const gfx_pedigree_desc* ped, gfx_schedule** out);
/* Same index arrays, no device upload: for inspecting the schedule (gfx_schedule_info / _copy) on a
 * machine without a GPU.  The handle must not be passed to a kernel entry point. */
int gfx_schedule_create_host(const gfx_pedigree_desc* ped, gfx_schedule** out);
void gfx_schedule_destroy(gfx_schedule* s);

typedef struct {
  int32_t n_rows;
  int32_t n_generations;   /* max generation + 1 */
  int32_t n_pairs;         /* distinct (sire, dam) of genotyped kids with both parents genotyped */
  int32_t n_pair_jobs;     /* sum over generations of pairs scheduled (a pair can run twice)     */
  int32_t n_singles;       /* genotyped animals that are in no partner pair                      */
  int32_t n_loopy;         /* rank blankets that are not the plain 2-generation tree             */
  int32_t n_blankets;
  int32_t max_blanket_nodes;
  int32_t max_width;       /* worst-case free-variable frontier over all blankets                */
  int32_t max_pairs_in_generation;
  int32_t n_trios;         /* genotyped animals with both parents genotyped                      */
  int32_t n_lone;          /* unpaired animals with neither ancestors nor genotyped children (never corrected) */
} gfx_schedule_info_t;

int gfx_schedule_info(const gfx_schedule* s, gfx_schedule_info_t* out);
/* Copies host-side index arrays out for inspection / tests.  which: 0 = pairs (n_pairs*2 rows:
 * sire_row, dam_row), 1 = pair jobs per generation offsets (n_generations+1), 2 = singles rows,
 * 3 = generation pair-job list (n_pair_jobs pair indices), 4 = rank flags (n_rows). */
int gfx_schedule_copy(const gfx_schedule* s, int which, int32_t* out_host, int64_t capacity);

/* ---------------------------------------------------------------------------------------------
 * 2-bit packing (replaces the uint8 DataFrame the reference keeps per worker,
 * correct_genotypes3.py:42-50; gfutils/gensrandaccess.py:38-50 yields the uint8 matrix)
 * codes: uint8 [n, ld_codes] with values 0,1,2 and 9 (anything > 2 packs as missing).
 * ------------------------------------------------------------------------------------------- */
int gfx_pack2(const uint8_t* codes_dev, int64_t n, int64_t s, int64_t ld_codes,
              uint32_t* packed_dev, int64_t wpr, void* stream);
int gfx_unpack2(const uint32_t* packed_dev, int64_t n, int64_t s, int64_t wpr,
                uint8_t* codes_dev, int64_t ld_codes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mendel rank ("peeling") kernel: CorrectGenotypes.mendelProbsSingle for every genotyped animal
 * and SNP (correct_genotypes3.py:85-166 with bayesnet/model.py:93-116 and the pgmpy query of
 * :119-121).  raw_dev receives the float16 score of :149-155 (1.0 where unobserved / lone).
 * ------------------------------------------------------------------------------------------- */
int gfx_mendel_rank(const gfx_schedule* s, const uint32_t* packed_dev, int64_t n_snps, int64_t wpr,
                    uint16_t* raw_dev, int64_t ld_raw, void* stream);

/* Fused form used by the pipeline: gfx_mendel_rank + gfx_rank_hist in ONE pass over the matrix (the scores are
 * written once and never re-read): raw_dev receives the scores with the 0xFFFF marker on missing cells, hist_dev
 * (65537 uint64, zeroed here) the histogram described below.  Rows of packed_dev must be padded with code 3. */
int gfx_mendel_rank_hist(const gfx_schedule* s, const uint32_t* packed_dev, int64_t n_snps, int64_t wpr,
                         uint16_t* raw_dev, int64_t ld_raw, uint64_t* hist_dev, void* stream);

/* Histogram of the raw float16 bit patterns over non-missing cells (65536 x uint64) plus the
 * number of missing cells in hist_dev[65536]: everything the rank transform and both quantiles need
 * (correct_genotypes3.py:593-606,666-667).  hist_dev must hold 65537 uint64 and is zeroed here.
 * Side effect: the score of every missing cell (always 1.0, :102) is replaced by the marker 0xFFFF,
 * which the correction kernels read as "E = 1" (:603). */
int gfx_rank_hist(uint16_t* raw_dev, int64_t ld_raw, const uint32_t* packed_dev, int64_t wpr,
                  int64_t n, int64_t n_snps, uint64_t* hist_dev, void* stream);

/* Per-animal sum over SNPs of E = elut[raw] (missing cells use missing_value):
 * error_checker.py:193-204 / preindex.py:259-266 individualSumProbs. */
int gfx_rank_rowsum(const uint16_t* raw_dev, int64_t ld_raw, const uint32_t* packed_dev, int64_t wpr,
                    int64_t n, int64_t n_snps, const double* elut_dev, double missing_value,
                    double* rowsum_dev, void* stream);

/* The same tally in the reference's own arithmetic (error_checker.py:193-204, preindex.py:259-266): the transformed
 * ranks are FLOAT16 there and DataFrame.sum(axis=1) adds the SNP columns one after the other into a float16
 * accumulator (every addition in float32, rounded to float16).  raw_dev: scores with the 0xFFFF marker on missing
 * cells (after gfx_mendel_rank_hist / gfx_rank_hist); e16_dev: 65536 float16 bit patterns, E of every score
 * pattern, entry 0xFFFF = value of a missing cell (-1 in error_checker.py:196-197, 1 in preindex.py:262-263);
 * rows_dev: n_sel row indices (the animals with both parents genotyped, error_checker.py:138-153) or NULL for all
 * n rows; rowsum16_dev: n_sel float16 bit patterns.  ld_raw must be a multiple of 64. */
int gfx_rank_rowsum_f16(const uint16_t* raw_dev, int64_t ld_raw, int64_t n, int64_t n_snps, const uint16_t* e16_dev,
                        const int32_t* rows_dev, int32_t n_sel, uint16_t* rowsum16_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Correction: correctPair + pair apply per generation (correct_genotypes3.py:169-327,684-848) and
 * correctSingle + single apply (:330-430,851-944).
 *   elut_dev     65536 doubles: transformed rank E for every raw bit pattern (built by the host
 *                with numpy so log() matches the reference bit for bit); missing cells are E = 1
 *   raw_dev      scores after gfx_rank_hist (missing cells marked 0xFFFF); packed_live is updated in place
 *   scratch_dev  scratch_words uint32, at least n_rows * (wpr + 32 * ceil(n_snps / 1024)) + 65552: the
 *                generation-start snapshot every worker of the reference reads (:687-689) -- neighbours are
 *                read from it, only the focal cell is live --, a fallback bitmap, and whatever is left for the
 *                lists of parked cells and deferred jobs (about 80 bytes per cell whose ancestry needs more
 *                than plain peeling; cells that do not fit are redone by a slower whole-cell kernel)
 *   post_dev     optional (may be NULL): posterior of every evaluated (job, side, SNP), 3 doubles at
 *                [((job * sides + side) * n_snps + snp) * 3]; cells that were not evaluated are left
 *                untouched (prefill with a sentinel)
 *   tallies_dev  int32 [2 * n_rows]: [0..n) corrections applied (n_errors), [n..2n) excluded by the
 *                rank gate (excludedNotBest); accumulated, not zeroed here
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  double test_p;        /* quantP_t, :666-667 */
  double suspect_t;     /* 0.2, :691 */
  double tiethreshold;
  double threshold;     /* threshold_pair or threshold_single */
  double err_thresh;
  const uint16_t* cutraw_dev; /* 65537 entries from gfx_build_cut_table, on the device */
  uint32_t tp_raw;            /* from gfx_build_cut_table */
  uint32_t reserved;
} gfx_correct_params;

/* HOST helper: E is monotone in the float16 score, so both rank comparisons of correctPair/correctSingle
 * become comparisons of raw bit patterns.  tp_raw = first pattern with E >= test_p (:181-185,342);
 * cutraw[r] = first pattern whose E > E(r) - suspect_t (:230,378), r = 65536 standing for E = 1.
 * elut_host: the 65536-entry table that is also uploaded as elut_dev. */
int gfx_build_cut_table(const double* elut_host, double test_p, double suspect_t, uint16_t* cutraw_host,
                        uint32_t* tp_raw);

int gfx_correct_pairs(const gfx_schedule* s, int32_t generation, uint32_t* packed_live_dev,
                      int64_t n_snps, int64_t wpr,
                      const uint16_t* raw_dev, int64_t ld_raw, const double* elut_dev,
                      const gfx_correct_params* params, uint32_t* scratch_dev, int64_t scratch_words,
                      double* post_dev, int32_t* tallies_dev, void* stream);

int gfx_correct_singles(const gfx_schedule* s, uint32_t* packed_live_dev,
                        int64_t n_snps, int64_t wpr,
                        const uint16_t* raw_dev, int64_t ld_raw, const double* elut_dev,
                        const gfx_correct_params* params, uint32_t* scratch_dev, int64_t scratch_words,
                        double* post_dev, int32_t* tallies_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The empirical (LD-window) blend of correctMatrix(empC != None).
 *
 * gfx_emp_index: the index of empirical/jalleledist3.py as device arrays -- per SNP column its window [wstart, wstart + wlen)
 * (jalleledist3.py:62-75, D11 at the chromosome end), the offset of its 3^wlen stored frequencies in freq_dev (state tuples in
 * itertools.product order, :31) and weight_empirical.  Windows of at most 9 SNPs.
 *
 * gfx_emp_rank_blend: correct_genotypes3.py:619-656 for every cell -- e_out[r, j] = mean(E[r, j], weight * (max - own) of the
 * normalised window frequencies of the animal's calls, getEmpProbs :69-82 / getCountTable jalleledist3.py:101-117) for observed
 * cells of SNPs whose window holds them, E[r, j] (1 for missing cells) otherwise; E = elut[raw].  A missing call inside the
 * window is summed over its completions (the reference raises there).
 *
 * gfx_emp_apply: the serial apply step of one generation of pairs (generation >= 0, :714-845) or of the singles (-1, :866-944)
 * with the blended decision (:746-779 / :883-908): post_dev holds the posteriors of the jobs (gfx_correct_pairs / _singles on a
 * scratch copy of the matrix, cells without a result = post_sentinel), snapshot_dev the matrix at the start of the generation,
 * live_dev the matrix being corrected (its window cells are the evidence).  One thread per focal animal, jobs in canonical
 * order, SNPs in column order.  tallies_dev as in gfx_correct_pairs.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const int32_t* wstart_dev;
  const int32_t* wlen_dev;
  const int64_t* foff_dev;
  const double* freq_dev;
  double weight;
  int32_t max_window;
  int32_t reserved;
} gfx_emp_index;

int gfx_emp_rank_blend(const uint32_t* packed_dev, int64_t n, int64_t n_snps, int64_t wpr, const uint16_t* raw_dev,
                       int64_t ld_raw, const double* elut_dev, const gfx_emp_index* index, double* e_out_dev, int64_t ld_e,
                       void* stream);
int gfx_emp_apply(const gfx_schedule* s, int32_t generation, uint32_t* live_dev, const uint32_t* snapshot_dev, int64_t n_snps,
                  int64_t wpr, const uint16_t* raw_dev, int64_t ld_raw, const double* elut_dev, const gfx_emp_index* index,
                  const gfx_correct_params* params, const double* post_dev, double post_sentinel, int32_t* tallies_dev,
                  void* stream);

/* Diagnostics: reads and clears 16 device counters (queries, parent-only answers, eliminations, big-table
 * reruns, nodes touched, suspect cells, job evaluations, children visited, ...) and switches counting
 * on (enable = 1) or off.  Counting is off by default; it serialises on atomics. */
int gfx_debug_counters(gfx_schedule* s, int enable, uint64_t* out16_host);

/* Which rank kernel the next gfx_mendel_rank / gfx_mendel_rank_hist calls on this schedule use: 0 = k_mendel_rank3 whenever the
 * score table is Mendel's and the rows are 16-byte aligned (default), 2 = k_mendel_rank, the leaner kernel when nearly every word
 * holds a missing call (the caller knows the missing fraction of the previous chunk from bin 65536 of its histogram).  Both
 * give the same scores and histogram bit for bit (correct_genotypes3.py:85-166). */
int gfx_rank_form(gfx_schedule* s, int form);

/* Diagnostics: times the dominant kernel of gfx_mendel_rank / gfx_mendel_rank_hist (k_mendel_rank3, or k_mendel_rank) by
 * itself -- two CUDA events recorded on the launching stream right around that launch, not around the histogram memset and
 * the loopy-row consumer of the same call.  gfx_rank_timing(s, 1) switches the recording on (one caller at a time),
 * gfx_rank_last_kernel_ms waits for the last recorded launch and returns its duration in milliseconds. */
int gfx_rank_timing(gfx_schedule* s, int enable);
int gfx_rank_last_kernel_ms(gfx_schedule* s, float* ms);

/* Device-side error flag raised by the inference kernels (e.g. GFX_E_TOO_WIDE); synchronises the stream.  The flag is sticky
 * per schedule: engines that share a schedule (lanes) share its one status word, so once any of them has seen the error every
 * later check fails as well, until gfx_reset_device_status (called at the start of a top-level correction call). */
int gfx_check_device_status(const gfx_schedule* s, void* stream);
int gfx_reset_device_status(gfx_schedule* s);

/* ---------------------------------------------------------------------------------------------
 * Empirical Mendelian-consistency scan (BASELINE config 5; trio selection of
 * pedigree/error_checker.py:138-153): for every genotyped animal with both parents genotyped,
 * the number of SNPs where all three calls are present (tested) and the number where
 * P(kid | sire, dam) = 0 under bayesnet/model.py:28-30 (inconsistent).  int64 [n_rows] each.
 * Rows must be 16-byte aligned (wpr % 4 == 0).  gfx_trio_scan zeroes the counters first; gfx_trio_scan_add
 * adds to them (a matrix held as column chunks: one call per chunk, SURVEY 8(d) tile rule).
 * ------------------------------------------------------------------------------------------- */
int gfx_trio_scan(const gfx_schedule* s, const uint32_t* packed_dev, int64_t n_snps, int64_t wpr,
                  int64_t* inconsistent_dev, int64_t* tested_dev, void* stream);
int gfx_trio_scan_add(const gfx_schedule* s, const uint32_t* packed_dev, int64_t n_snps, int64_t wpr,
                  int64_t* inconsistent_dev, int64_t* tested_dev, void* stream);

/* Founders (gfutils/extract_founders.py:106-110): HOST call.  restrict_host is NULL or n_ped flags;
 * writes pedigree indices of animals lacking a sire or a dam inside the (restricted) pedigree. */
int gfx_founders(const gfx_pedigree_desc* ped, const uint8_t* restrict_host, int32_t* out_idx_host,
                 int32_t* out_n);

/* ---------------------------------------------------------------------------------------------
 * LD-window joint genotype counts (empirical/jalleledist3.py:120-141 countJointFrq, the integer
 * part): for each of n_windows windows [win_start[i], win_start[i]+win_len[i]) count, over rows
 * whose window cells all pass the mask (mask bit = 1, 1 bit per SNP, `mwpr` uint32 words per row,
 * may be NULL = all pass) and are non-missing, the rows matching each of the 3^win_len state tuples
 * (first SNP of the window is the most significant base-3 digit, itertools.product order).
 * The nested inner windows of :133-138 (cells i .. len-1-i, i = 1 .. int((len-2)/2) - 1) are counted
 * separately over the same rows -- a row with a missing call in an OUTER cell still counts there, as in the
 * reference -- and follow the full histogram: [3^len][3^(len-2)][3^(len-4)]...
 * counts_dev: uint32 [n_windows * stride], stride >= 3^max_len + 3^(max_len-2) + ... (the counters above).
 * ------------------------------------------------------------------------------------------- */
int gfx_window_counts(const uint32_t* packed_dev, const uint32_t* mask_dev, int64_t n, int64_t n_snps,
                      int64_t wpr, int64_t mwpr, const int32_t* win_start_dev, const int32_t* win_len_dev,
                      int32_t n_windows, int32_t stride, uint32_t* counts_dev, void* stream);

/* Host-side inference on one blanket (the same routine the kernels run), for tests and for
 * building tables: nodes in topological order, p1/p2 local parent indices or -1, code 0..2 or 3
 * (unobserved).  out = normalised P(query | evidence) (NaN if P(evidence) = 0). */
int gfx_infer_host(int32_t n, const int8_t* p1, const int8_t* p2, const uint8_t* code, int32_t query,
                   double* out3);
/* Same for the joint=False query of two animals (correctPair): out6 = P(q0|e) then P(q1|e); both are
 * NaN when either animal's component has P(evidence) = 0, like pgmpy's normalised joint. */
int gfx_infer_pair_host(int32_t n, const int8_t* p1, const int8_t* p2, const uint8_t* code, int32_t q0,
                        int32_t q1, double* out6);
/* Host-side children term (bayesnet/model.py:65-89 with numpy's summation order). kid/par: n codes,
 * par 0..2 or anything else = unknown. */
int gfx_kids_term_host(int32_t n, const uint8_t* kid, const uint8_t* par, double* out3);

/* ---------------------------------------------------------------------------------------------
 * Simulator pieces on the device (simulate.py:187-340 gene drop + error injection, :426-515 evaluation;
 * quickgsim/genome.py:86-115 meiosis: gfx_gene_drop reduces it to independent loci, i.e. free recombination, gfx_gene_drop_linked
 * follows it crossover by crossover).
 *   gfx_gene_drop      hap_dev uint32 [n, wpr]: bit 2i = paternal, bit 2i+1 = maternal allele of SNP i.  Rows are processed
 *                      level by level: order_dev lists the rows, level_off_host[l..l+1] delimits level l, the parents of a
 *                      level are in earlier levels; sire_row / dam_row = -1 for founders (alleles with frequency 1/2).
 *   gfx_haps_to_codes  2-bit genotype codes (padding = 3) from the haplotypes.
 *   gfx_inject_errors  every observed cell with probability `rate` shows one of the two other genotypes.
 *   gfx_confusion      out8_dev: errors, fixed, blanked, wrong fix, new errors, missed, cells, 0.
 * ------------------------------------------------------------------------------------------- */
int gfx_gene_drop(int32_t n, const int32_t* sire_row_dev, const int32_t* dam_row_dev, const int32_t* order_dev,
                  const int32_t* level_off_host, int32_t n_levels, int64_t n_snps, int64_t wpr, uint64_t seed,
                  uint32_t* hap_dev, void* stream);
/* The same drop with linkage: quickgsim's meiosis (quickgsim/genome.py:74-115).  SNPs are in map order, chromosome c holds the
 * SNPs chrom_start[c] .. chrom_start[c + 1] - 1; per gamete and chromosome max(1, Poisson(morgans[c])) crossovers on distinct
 * interior markers drawn with the probabilities whose running sum per chromosome is cum_p (Chrom.cm_to_p, :35-41: proportional
 * to the genetic distance to the previous marker; 0 at the first marker, 1 from the last but one on), random start strand,
 * strands alternate at the crossover markers.  All map arrays on the device. */
int gfx_gene_drop_linked(int32_t n, const int32_t* sire_row_dev, const int32_t* dam_row_dev, const int32_t* order_dev,
                         const int32_t* level_off_host, int32_t n_levels, int64_t n_snps, int64_t wpr, uint64_t seed,
                         int32_t n_chrom, const int32_t* chrom_start_dev, const float* morgans_dev, const float* cum_p_dev,
                         uint32_t* hap_dev, void* stream);
int gfx_haps_to_codes(const uint32_t* hap_dev, int64_t n, int64_t n_snps, int64_t wpr, uint32_t* packed_dev, void* stream);
int gfx_inject_errors(uint32_t* packed_dev, int64_t n, int64_t n_snps, int64_t wpr, double rate, uint64_t seed, void* stream);
int gfx_confusion(const uint32_t* truth_dev, const uint32_t* obs_dev, const uint32_t* fixed_dev, int64_t n, int64_t n_snps,
                  int64_t wpr, uint64_t* out8_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
