//go:build cgo

// Package gpu is the cgo binding a LexicMap maintainer would add (lexicmap/cmd/gpu/lexicmap_gpu.go) to route
// `lexicmap search` through liblexicmap_gpu.so. SOURCE ONLY: this image has no Go toolchain, so it is not compiled or
// tested here; the C++ driver csrc/search_cli.cpp exercises the same C ABI.
//
// It replaces, for the search hot path only:
//     NewIndexSearcher  lexicmap/cmd/lib-index-search.go:237   ->  gpu.Open
//     (*Index).Search   lexicmap/cmd/lib-index-search.go:1191  ->  (*Index).SearchBatch
package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../lexicmap_b200 -llexicmap_gpu
#include <stdlib.h>
#include "lexicmap_gpu.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// Params mirrors IndexSearchingOptions + SeqComparatorOptions + Chaining2Options.
type Params = C.lmg_params

// HSP is one TSV row (0-based inclusive coordinates).
type HSP = C.lmg_hsp

type Index struct{ h *C.lmg_index }

func lastErr() error { return errors.New(C.GoString(C.lmg_last_error())) }

// DefaultParams returns the `lexicmap search` flag defaults.
func DefaultParams() Params { var p Params; C.lmg_default_params(&p); return p }

// Open loads the .lmi directory into the HBM of one GPU.
func Open(dir string, device, shard, nShards int) (*Index, error) {
	cs := C.CString(dir)
	defer C.free(unsafe.Pointer(cs))
	var h *C.lmg_index
	if C.lmg_index_open(cs, C.int(device), C.int(shard), C.int(nShards), &h) != 0 {
		return nil, lastErr()
	}
	return &Index{h}, nil
}

func (idx *Index) Close() { C.lmg_index_close(idx.h) }

// GenomeName maps SearchResult.BatchGenomeIndex to the genome ID (idx.BatchGenomeIndex2GenomeID).
func (idx *Index) GenomeName(bgi uint64) string {
	var s *C.char
	C.lmg_genome_name(idx.h, C.uint64_t(bgi), &s)
	return C.GoString(s)
}

// Results owns the rows of one batch until Free is called (cf. idx.RecycleSearchResults).
type Results struct {
	r    *C.lmg_results
	Rows []HSP
	pool []byte
}

func (r *Results) SeqID(i int) string {
	var s *C.char
	C.lmg_results_seq_id(r.r, C.uint64_t(i), &s)
	return C.GoString(s)
}
func (r *Results) CIGAR(i int) string {
	h := r.Rows[i]
	return string(r.pool[h.cigar_off : uint64(h.cigar_off)+uint64(h.cigar_len)])
}
func (r *Results) Free() { C.lmg_results_free(r.r) }

// SearchBatch runs Index.Search for a batch of upper- or lower-case query sequences.
func (idx *Index) SearchBatch(p *Params, seqs [][]byte) (*Results, error) {
	off := make([]C.uint64_t, len(seqs)+1)
	total := 0
	for i, s := range seqs {
		total += len(s)
		off[i+1] = C.uint64_t(total)
	}
	buf := make([]byte, 0, total+16)
	for _, s := range seqs {
		buf = append(buf, s...)
	}
	buf = append(buf, make([]byte, 16)...)
	var r *C.lmg_results
	if C.lmg_search_batch(idx.h, p, (*C.uint8_t)(unsafe.Pointer(&buf[0])), &off[0], C.int32_t(len(seqs)), &r) != 0 {
		return nil, lastErr()
	}
	var rows *C.lmg_hsp
	var n, pl C.uint64_t
	var pool *C.char
	C.lmg_results_rows(r, &rows, &n, &pool, &pl)
	res := &Results{r: r}
	if n > 0 {
		res.Rows = unsafe.Slice((*HSP)(unsafe.Pointer(rows)), int(n))
	}
	if pl > 0 {
		res.pool = unsafe.Slice((*byte)(unsafe.Pointer(pool)), int(pl))
	}
	return res, nil
}
