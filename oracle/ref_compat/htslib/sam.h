// Stand-in for <htslib/sam.h> (+ the parts of <htslib/hts.h> it pulls in): exactly the types, macros and functions the
// reference's src/bamsignals.cpp uses, with htslib >= 1.10's names, field names, types and argument order, so that the
// reference file compiles UNCHANGED into oracle/_ref/ (see oracle/Makefile).
// TEST INFRASTRUCTURE.  htslib itself is an un-vendored dependency of the reference (LinkingTo: Rhtslib >= 1.13.1,
// DESCRIPTION:30; no lock file) and is absent offline; ../hts_compat.cpp implements these entry points over zlib from
// the SAM/BAM specification and htslib's documented iterator behaviour.
// Call sites in the reference: sam_open :202, bam_index_load :207, hts_idx_destroy :217, sam_close :218,
// sam_hdr_read :95, bam_name2id :27, bam_hdr_destroy :134, bam_init1 :250, bam_itr_queryi :267, bam_itr_next :271,
// bam_itr_destroy :287, bam_destroy1 :290, bam_endpos :17, BAM_FREVERSE :12; the fixture writer (:496-534) also names
// sam_hdr_write, sam_read1, bam_write1 and bam_index_build.
#pragma once
#include <stdint.h>

#include "bgzf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;                     // htslib >= 1.10

// ---- hts.h ---------------------------------------------------------------------------------------------------------
typedef struct htsFile {
    uint32_t is_bin : 1, is_write : 1, is_be : 1, is_cram : 1, is_bgzf : 1, dummy : 27;
    int64_t lineno;
    char* fn;
    union {
        BGZF* bgzf;
        void* cram;
        void* hfile;
    } fp;
} htsFile;

typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;

void hts_idx_destroy(hts_idx_t* idx);
void hts_itr_destroy(hts_itr_t* iter);

// ---- sam.h ---------------------------------------------------------------------------------------------------------
typedef htsFile samFile;

typedef struct sam_hdr_t {
    int32_t n_targets, ignore_sam_err;
    uint32_t l_text;
    uint32_t* target_len;
    char** target_name;
    char* text;
    void* sdict;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t* data;
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy : 2, : 30;
} bam1_t;

#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)   // bit 1: consumes query; bit 2: consumes reference

samFile* hts_open(const char* fn, const char* mode);
int hts_close(htsFile* fp);
#define sam_open(fn, mode) (hts_open((fn), (mode)))
#define sam_close(fp) hts_close(fp)

sam_hdr_t* sam_hdr_read(samFile* fp);
int sam_hdr_write(samFile* fp, const sam_hdr_t* h);
void sam_hdr_destroy(sam_hdr_t* h);
#define bam_hdr_destroy(h) sam_hdr_destroy(h)
int sam_hdr_name2tid(sam_hdr_t* h, const char* ref);
#define bam_name2id(h, ref) sam_hdr_name2tid((h), (ref))

bam1_t* bam_init1(void);
void bam_destroy1(bam1_t* b);
hts_pos_t bam_endpos(const bam1_t* b);

hts_idx_t* sam_index_load(htsFile* fp, const char* fn);
hts_idx_t* hts_idx_load(const char* fn, int fmt);
#define HTS_FMT_CSI 0
#define HTS_FMT_BAI 1
#define bam_index_load(fn) hts_idx_load((fn), HTS_FMT_BAI)
int sam_index_build(const char* fn, int min_shift);
#define bam_index_build(fn, min_shift) (sam_index_build((fn), (min_shift)))

hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(htsFile* htsfp, hts_itr_t* itr, bam1_t* r);
#define bam_itr_destroy(iter) hts_itr_destroy(iter)
#define bam_itr_queryi(idx, tid, beg, end) sam_itr_queryi(idx, tid, beg, end)
#define bam_itr_next(htsfp, itr, r) sam_itr_next((htsfp), (itr), (r))

int sam_read1(samFile* fp, sam_hdr_t* h, bam1_t* b);
int bam_write1(BGZF* fp, const bam1_t* b);

#ifdef __cplusplus
}
#endif
