/*
 * phyloaln_b200.h -- C ABI of libphyloaln_b200.so, the B200 (sm_100a) replacement for
 * the search hot path of PhyloAln.
 *
 * Plain C, caller-owned buffers, integer return codes (0 = ok, <0 = error; message via
 * pa_last_error), no exceptions and no exit() across the boundary.  One pa_ctx per GPU;
 * calls on one ctx must be serialised by the caller; contexts are independent.
 *
 * What each entry point replaces in the reference (/root/reference):
 *   pa_add_hmm / pa_commit_hmms  the `ref_hmm/{g}.hmm` argument of the hmmsearch command line
 *                                (lib/aln.py:664-666); the ASCII file is read by pa_hmm_file_parse (below),
 *                                its arrays cross the ABI here.  Profiles of 1..8192 columns.
 *   pa_load_reads                raw2target(): six-frame / codon translation or reverse
 *                                complement of a chunk of reads (lib/aln.py:233-259) fused with
 *                                HMMER's digitisation of all.target.fasta.
 *   pa_load_proteins             hmmsearch reading an already-protein target file.
 *   pa_search                    the `hmmsearch` child process per alignment
 *                                (lib/aln.py:664-667): MSV/SSV -> bias -> Viterbi -> Forward ->
 *                                domain definition -> optimal-accuracy alignment, for ALL
 *                                loaded profiles against the loaded target chunk.
 *   pa_get_hits                  the information extract_reads() pulls out of the hmmsearch
 *                                text report via Bio.SearchIO (lib/aln.py:511-558).
 *   pa_map_hits / pa_map_hits_ex pretreat_hits() (short-hit drop, -b outgroup pre-filter, terminal
 *                                trimming; lib/assemble.py:767-873), read_hit.__init__'s column walk and
 *                                --extend_pos extension (:13-306) and read_hit.consensus()'s per-column
 *                                accumulation (:309-633).
 *   pa_get_hit_pp                the PP annotation lines of the same report (Bio.SearchIO aln_annotation).
 *   pa_hmm_file_parse            hmmsearch's reader of ref_hmm/{g}.hmm (HMMER3/f ASCII, written by hmmbuild,
 *                                lib/aln.py:66-76).
 *   pa_fasta_*                   lib.read_fastx(..., 'fasta') on all.raw.fasta (lib/library.py:186-198).
 *   pa_ingest_*                  prepare_target()'s read / tag / split / merge pass (lib/aln.py:274-433).
 *   pa_score_all                 (no reference equivalent) raw filter scores for every
 *                                (profile, target) pair: E-value calibration and parity tests.
 */
#ifndef PHYLOALN_B200_H
#define PHYLOALN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_ABI_VERSION 3
#define PA_AMINO 0
#define PA_DNA 1

/* moltype of pa_load_reads */
#define PA_TRANSLATE 0 /* prot / codon / dna_codon modes: translate frames            */
#define PA_NUCLEOTIDE 1 /* dna modes: forward (+ reverse complement) nucleotide targets */

typedef struct pa_ctx pa_ctx;

typedef struct {
  double F1, F2, F3; /* MSV, Viterbi, Forward filter P-value thresholds (0.02, 1e-3, 1e-5) */
  int do_bias;       /* bias filter on (default 1; --nobias -> 0)                          */
  int do_null2;      /* null2 score correction (default 1; --nonull2 -> 0)                 */
  int do_max;        /* --max: all filters off                                             */
  int reserved;
} pa_opts;

typedef struct {
  int32_t hmm;    /* profile index (order of pa_add_hmm)                                   */
  int32_t target; /* target index in the loaded chunk: read*nframes + frame               */
  int32_t ndom_in_seq, dom_idx;
  int32_t ienv, jenv, iali, jali, hmmfrom, hmmto; /* 1-based inclusive                      */
  float envsc, domcorrection, oasc, dombias, dom_bits;
  double dom_lnP;
  float seq_bits, pre_bits, seqbias, fwdsc, nullsc;
  double seq_lnP;
  int32_t ali_off, ali_len; /* into the model/target string pools of pa_get_hits            */
  int32_t multidomain_region;
} pa_hit;

typedef struct {
  uint64_t n_pairs;    /* (profile, target) pairs entering the MSV stage                    */
  uint64_t n_cells;    /* sum of L*M over those pairs                                       */
  uint64_t n_past_ssv; /* pairs whose SSV score required the full MSV recurrence            */
  uint64_t n_past_msv, n_past_bias, n_past_vit, n_past_fwd;
  uint64_t n_domains;
  uint64_t kernel_launches;                                /* CUDA kernels launched so far  */
  float ms_translate, ms_msv, ms_bias, ms_vit, ms_fwd, ms_domain; /* last call, CUDA events */
  float ms_h2d, ms_total;
  float ms_domain_kernel;                                  /* domain_kernel alone (ms_domain adds D2H + host scoring) */
  float vit_exact_pairs;                                   /* pairs the exact Viterbi kernel had to run after the 16-bit upper-bound pass */
} pa_stats;

/* per-hit column mapping result (pa_map_hits) */
typedef struct {
  int32_t keep;                 /* 0 = dropped by pretreat_hits                             */
  int32_t qstart, qend;         /* reference columns after trimming                         */
  int32_t ali_from, ali_to;     /* target coordinates after trimming (residue units)        */
  int32_t start_trim, end_trim;
  int32_t read_from, read_to;   /* 1-based nucleotide interval on the raw read              */
  int32_t strand;               /* +1 / -1                                                  */
  int32_t col_off, ncols;       /* into base/codon pools: ncols = qend-qstart+1             */
  int32_t unmap_off, n_unmap;   /* into the unmapped-position pool                          */
  int32_t ext_start, ext_end;   /* columns added by --extend_pos (already inside qstart/qend/ncols) */
} pa_maprec;

/* options of pa_map_hits_ex = the arguments of pretreat_hits() / read_hit() (lib/assemble.py:767, :13) */
typedef struct {
  int32_t trim_pos;        /* --trim_pos (5)                                                  */
  int32_t extend_pos;      /* --extend_pos (0 = off)                                          */
  int32_t gencode;         /* --gencode: translation of the flanking codons during extension  */
  int32_t prefilter;       /* pretreat_hits' `outgroup_score and no_assemble` pre-filter (-b) */
  double pos_freq;         /* --pos_freq (0.2)                                                */
  double outgroup_weight;  /* --outgroup_weight (1)                                           */
} pa_map_opts;

/* optional accumulation request of pa_map_hits_ex = read_hit.consensus() (lib/assemble.py:309-633):
 * kept hit h adds rep[h] to hist[(bin_row[bin[h]] + col - 1)*96 + k*32 + sym] for every reference column
 * col it covers (k = codon position 0..2, or 0 in nucleotide mode; sym as in `comp`), records the smallest
 * contributing hit index in `first` (same shape, INT32_MAX = none; Python dict insertion order) and widens
 * qrange[2*bin] / qrange[2*bin+1] (min qstart / max qend; caller initialises). bin[h] < 0 = skip. */
typedef struct {
  const int32_t *bin, *rep;
  int32_t nbins;
  const int64_t *bin_row; /* nbins + 1 */
  int32_t *hist, *first;  /* bin_row[nbins] * 96 each, host memory, overwritten */
  int32_t *qrange;        /* nbins * 2, host memory, read and updated           */
} pa_accum;

int pa_abi_version(void);
pa_ctx *pa_create(int device, char *err, int errlen);
void pa_destroy(pa_ctx *ctx);
const char *pa_last_error(pa_ctx *ctx);
int pa_set_opts(pa_ctx *ctx, const pa_opts *opts);

/* Profiles. Arrays are the numbers of the HMMER3/f file: -ln(p), +inf for '*'.
 * mat, ins: (M+1)*K row-major (row 0 = node-0 line; mat row 0 ignored); t: (M+1)*7 in file order
 * MM MI MD IM II DM DD; compo: K values or NULL; cons: M consensus letters or NULL;
 * evparam: {MSV mu, lambda, VITERBI mu, lambda, FORWARD tau, lambda}. Returns profile index. */
int pa_add_hmm(pa_ctx *ctx, const char *name, int abc, int M, const double *mat, const double *ins,
               const double *t, const double *compo, const char *cons, const float *evparam);
int pa_set_evparam(pa_ctx *ctx, int hmm, const float *evparam);
int pa_commit_hmms(pa_ctx *ctx);
/* the same profile set on n GPUs: profiles (and options) of ctxs[0] are configured once and uploaded to every context */
int pa_commit_hmms_multi(pa_ctx **ctxs, int n);
int pa_num_hmms(pa_ctx *ctx);

/* Targets. nt: concatenated ASCII nucleotides, off[nreads+1] byte offsets (HOST memory).
 * Produces nreads*nframes targets: TRANSLATE: frames pos1..3 (+ rev_pos1..3 unless no_reverse;
 * only pos1 if codon_only); NUCLEOTIDE: read (+ reverse complement). Returns #targets or <0.
 * -2: invalid nucleotide character (Biopython would raise TranslationError) or unsupported genetic code;
 * -3: capacity (2*nt + 32*reads >= 4 GiB, or more than 65,535 distinct target lengths): split the chunk and retry. */
int64_t pa_load_reads(pa_ctx *ctx, const char *nt, const uint64_t *off, uint32_t nreads, int moltype,
                      int gencode, int no_reverse, int codon_only);
int64_t pa_load_proteins(pa_ctx *ctx, const char *aa, const uint64_t *off, uint32_t nseq, int abc);
int pa_num_frames(pa_ctx *ctx);
/* copy the device-resident target residues back as ASCII: out_off[ntargets+1], out (cap bytes) */
int64_t pa_get_targets(pa_ctx *ctx, char *out, uint64_t cap, uint64_t *out_off);

/* Search every committed profile against the loaded targets. Returns #hits (domains) or <0. */
int64_t pa_search(pa_ctx *ctx);
int64_t pa_get_hits(pa_ctx *ctx, pa_hit *hits, int64_t cap, char *model_pool, char *target_pool,
                    int64_t poolcap);
int64_t pa_hit_pool_bytes(pa_ctx *ctx); /* bytes needed for each string pool of pa_get_hits */
/* posterior-probability annotation of every alignment (the "PP" line of the hmmsearch text report: '0'..'9', '*',
 * '.' for deletes), same offsets/lengths as the pools of pa_get_hits. Returns bytes written or <0. */
int64_t pa_get_hit_pp(pa_ctx *ctx, char *pp_pool, int64_t poolcap);
int pa_get_stats(pa_ctx *ctx, pa_stats *st);

/* Raw filter scores (nats) of profile `hmm` for all loaded targets; any output may be NULL.
 * msv_raw: final xJ byte (-1 overflow); vit_raw: final xC word (32767 overflow). */
int pa_score_all(pa_ctx *ctx, int hmm, float *msv, int32_t *msv_raw, float *vit, int32_t *vit_raw,
                 float *fwd, float *bias);

/* Measured issue rates of the instructions the filters are built from (no reference equivalent;
 * roofline denominators). mode 0 VIADDMNMX.S16x2.RELU, 1 VIMNMX3.S16x2, 2 DPX SSV mix (8 VIADDMNMX +
 * 4 VIMNMX3), 3 VIADDMNMX.S32, 4 HFMA2.RELU, 5 fp16 SSV mix of msv_tmem_kernel (8 HFMA2.RELU + 4
 * VIMNMX3.S16x2 per 16 cells), 6 HFMA2.RELU warps next to DPX-mix warps (even / odd warps of a CTA), 7 both formulations in
 * every warp; 6 and 7 count the cell-update instructions only (2 cells each). Result in giga thread-instructions per second
 * over the whole GPU. */
int pa_microbench_dpx(pa_ctx *ctx, int mode, double *ginstr_per_s);

/* The order in which the MSV kernels walk their work items (no reference equivalent): item -> (profile slot, target tile),
 * profile-major inside an L2-sized super-tile of `super_tiles` target tiles, super-tile-major overall. Host evaluation of the
 * device function, for tests. 0 ok, -2 bad argument. */
int pa_msv_item_decode(uint64_t item, uint32_t nlist, uint32_t ntiles, uint32_t super_tiles, uint32_t *slot, uint32_t *tile);

/* cudaProfilerStart / cudaProfilerStop (on != 0 / on == 0): lets `ncu --profile-from-start off` capture
 * only the timed region of bench.py. No reference equivalent. */
int pa_profiler_range(int on);

/* Hit -> reference-column mapping (extend_pos = 0). For hit h: model/aligned strings (PhyloAln
 * form: upper case, '-' for gaps) at ali_off[h]..ali_off[h+1]; raw read nucleotides at
 * read_off[h]..; hmm_from/to, ali_from/to 1-based; frame 0..2 or -1 (nucleotide mode);
 * strand +1 / -1 ('rev'). site_freq: per hit pointer index into comp tables:
 * comp[(comp_off[h] + (col-1))*32 + sym] = count of symbol `sym` (A=0..Z=25, '-'=26,'*'=27)
 * in reference column col; outputs rec[nhits] and pools (base per column, 3 nt per column,
 * unmapped 1-based read positions). */
int pa_map_hits(pa_ctx *ctx, int32_t nhits, const char *model, const char *aligned, const int64_t *ali_off,
                const char *reads, const int64_t *read_off, const int32_t *hmm_from, const int32_t *hmm_to,
                const int32_t *ali_from, const int32_t *ali_to, const int32_t *frame, const int32_t *strand,
                const int32_t *comp, const int64_t *comp_off, const int32_t *ncols_ref, int trim_pos,
                double pos_freq, pa_maprec *rec, char *base_pool, char *codon_pool, int32_t *unmap_pool,
                int64_t pool_cols_cap, int64_t unmap_cap);
/* Same with every option of the reference: --extend_pos extension of untrimmed ends (lib/assemble.py:82-150,
 * 203-298), the -b outgroup pre-filter (lib/assemble.py:771-798; out_score = get_ref_score()'s outgroup_score as
 * rows of ncols_ref[h] doubles starting at out_score[out_off[h]], n_out[h] rows incl. PhyloAln_Alt, NaN = the
 * sequence has no score in that column; may be NULL), and the consensus accumulation (acc may be NULL).
 * Pools need sum(hmm_to-hmm_from+2 + 2*extend_pos) columns. */
int pa_map_hits_ex(pa_ctx *ctx, int32_t nhits, const char *model, const char *aligned, const int64_t *ali_off,
                   const char *reads, const int64_t *read_off, const int32_t *hmm_from, const int32_t *hmm_to,
                   const int32_t *ali_from, const int32_t *ali_to, const int32_t *frame, const int32_t *strand,
                   const int32_t *comp, const int64_t *comp_off, const int32_t *ncols_ref, const double *out_score,
                   const int64_t *out_off, const int32_t *n_out, const pa_map_opts *opts, const pa_accum *acc,
                   pa_maprec *rec, char *base_pool, char *codon_pool, int32_t *unmap_pool, int64_t pool_cols_cap,
                   int64_t unmap_cap);

/* merge_hits() of the reference (lib/assemble.py:889-954), histogram part, as a segmented left fold: chain j merges items
 * chain_off[j] .. chain_off[j+1]-1 in order (newhit = merge_hits(newhit, hit) repeatedly, lib/assemble.py:995-1766).  Item i
 * carries the counts of reference columns item_q0[i] .. item_q0[i]+item_ncols[i]-1 as [column][3][32] ints at row item_row[i] of
 * hist_in (layout of pa_accum.hist) and, in key_in, the insertion rank of each symbol inside the item's own dictionaries.
 * Per chain (columns 1 .. chain_ncols[j] at row chain_row[j]): hist_out = summed counts; who_out = position in the chain of the
 * first item with a non-zero count (INT32_MAX = none) and key_out = that item's rank -- together the Python dict insertion
 * order of the merged hit.  qend / seqids / count / fill_gap bookkeeping stays on the host (phyloaln_b200.mapping). */
int pa_merge_hists(pa_ctx *ctx, int32_t nchains, const int32_t *chain_off, const int32_t *item_q0, const int32_t *item_ncols,
                   const int64_t *item_row, const int32_t *hist_in, const int32_t *key_in, const int64_t *chain_row,
                   const int32_t *chain_ncols, int32_t *hist_out, int32_t *who_out, int32_t *key_out);

/* ---- native target preparation (host only; replaces the body of prepare_target(), lib/aln.py:274-433,
 * and the text iteration of read_fastx(return_iter=True), lib/library.py:157-176).  Files are inflated and
 * parsed in parallel; tagging, sliding-window splitting, exact-duplicate merging and the 10 MB chunk order are
 * reproduced byte for byte (all.raw.fasta, total_counts.tsv, merged_redundance.txt).  Test knob: the environment
 * variable PA_INGEST_EACH_SIZE replaces the 10,000,000-byte chunk size (lib/aln.py:283) so that the >1000-chunk fold can
 * be compared with the reference on small inputs; leave it unset in production. ---- */
typedef struct pa_ingest pa_ingest;
pa_ingest *pa_ingest_create(int merge_len, int split_len /* 0 = off */, int split_slide /* 0 = split_len/2 */, int nthreads);
void pa_ingest_destroy(pa_ingest *ing);
const char *pa_ingest_last_error(pa_ingest *ing);
/* files in the order of the reference's loops: for species, for file j (file_index = j + 1); format guess|fastq|fasta */
int pa_ingest_add_file(pa_ingest *ing, const char *path, const char *species, int file_index, const char *format);
int pa_ingest_run(pa_ingest *ing);            /* = pa_ingest_start + pa_ingest_wait */
int64_t pa_ingest_num_records(pa_ingest *ing);
/* Streaming: pa_ingest_start returns at once; inflate/parse threads and the single order-dependent pass run in the
 * background (at most nthreads files are inflated ahead of the pass).  Records become visible in EMISSION order
 * (= the order of all.raw.fasta unless the reference's >1000-chunk fold happened, see pa_ingest_final_order) as soon as
 * they are final, so the GPU search of the first records overlaps the decoding of the rest.  pa_ingest_emitted: records
 * visible so far (monotone; any thread); pa_ingest_done: 1 once the pass has ended; pa_ingest_wait: join, 0 or the error. */
int pa_ingest_start(pa_ingest *ing);
int pa_ingest_wait(pa_ingest *ing);
int pa_ingest_done(pa_ingest *ing);
int64_t pa_ingest_emitted(pa_ingest *ing);
/* like pa_ingest_pack, for emitted records [first, first+count) while the ingest may still be running */
int64_t pa_ingest_pack_emitted(pa_ingest *ing, uint64_t first, uint64_t count, char *nt_out, uint64_t cap, uint64_t *off_out);
int64_t pa_ingest_emitted_bytes(pa_ingest *ing, uint64_t first, uint64_t count);      /* sum of their sequence lengths */
/* ids of the records idx[0..n) (emission indices), concatenated into out (may be NULL: sizes only); out_off[n+1] */
int64_t pa_ingest_emitted_ids(pa_ingest *ing, const uint64_t *idx, uint64_t n, char *out, uint64_t cap, uint64_t *out_off);
int64_t pa_ingest_emitted_seqs(pa_ingest *ing, const uint64_t *idx, uint64_t n, char *out, uint64_t cap, uint64_t *out_off); /* same, sequences */
/* after the end: number of records whose id repeats an earlier record's id (the reference's read_fastx dictionary,
 * lib/library.py:186-198, keeps one entry per id: first position, last sequence; 0 = every record is its own target).
 * Counted on 64-bit id hashes: 0 is exact, a non-zero count tells the caller to compare the ids themselves. */
int64_t pa_ingest_repeated_ids(pa_ingest *ing);
/* after the end: 0 = all.raw.fasta order is the emission order; 1 = the chunk fold of check_file_num (lib/aln.py:262-271)
 * permuted it and emitted_index[final position] (num_records entries, may be NULL) is filled */
int pa_ingest_final_order(pa_ingest *ing, uint64_t *emitted_index);
/* borrowed views, valid until pa_ingest_destroy: the all.raw.fasta text, its record table (offsets/lengths of ids and
 * sequences), the total_counts.tsv and merged_redundance.txt texts. Any pointer may be NULL. */
int pa_ingest_views(pa_ingest *ing, const char **raw, uint64_t *raw_len, const uint64_t **id_off, const uint64_t **id_len,
                    const uint64_t **seq_off, const uint64_t **seq_len, const char **counts, uint64_t *counts_len,
                    const char **merged, uint64_t *merged_len);
int pa_ingest_write(pa_ingest *ing, const char *raw_fasta, const char *total_counts, const char *merged_redundance);
/* sequences [first, first+count) of the final (all.raw.fasta) order concatenated + offsets (count+1), ready for pa_load_reads */
int64_t pa_ingest_pack(pa_ingest *ing, uint64_t first, uint64_t count, char *nt_out, uint64_t cap, uint64_t *off_out);

/* ---- HMMER3/[a-f] ASCII reader (host only): the numbers of ref_hmm/{g}.hmm (first model of the file), ready for pa_add_hmm.
 * Call once with mat == NULL for M / abc / name / stats, then with arrays: mat, ins (M+1)*K, t (M+1)*7 (-ln p, +inf for '*'),
 * compo K (NaN when absent), cons / rf M+1 bytes (empty string when absent), map M ints (map[0] = -1 when absent),
 * stats {MSV mu, lambda, VITERBI mu, lambda, FORWARD tau, lambda} (NaN when absent). 0 ok, <0 malformed / unreadable. ---- */
int pa_hmm_file_parse(const char *path, int32_t *M, int32_t *abc, char *name, int namecap, double *mat, double *ins, double *t,
                      double *compo, char *cons, char *rf, int32_t *map, float *stats, int32_t *nseq, double *effn);

/* ---- all.raw.fasta loader (host only): what dropin.hmmsearch_step needs from the file prepare_target() wrote --
 * ids, concatenated nucleotides and offsets ready for pa_load_reads -- with the dictionary semantics of the reference's
 * read_fastx(..., 'fasta') (lib/library.py:186-198): id = header up to the first blank, a repeated id keeps its first
 * position and its last sequence. gz or plain. Views are borrowed until pa_fasta_close. ---- */
typedef struct pa_fasta pa_fasta;
pa_fasta *pa_fasta_open(const char *path, char *err, int errlen);
int64_t pa_fasta_num_records(pa_fasta *f);
int pa_fasta_views(pa_fasta *f, const char **ids, const uint64_t **id_off /* n+1 */, const char **nt, const uint64_t **off /* n+1 */);
void pa_fasta_close(pa_fasta *f);

#ifdef __cplusplus
}
#endif
#endif
