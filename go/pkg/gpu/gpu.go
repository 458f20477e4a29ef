// Package gpu binds libquivergpu.so (include/quiver_gpu.h) and exposes a core.Index /
// hybrid.Index implementation backed by one B200.
//
// NOTE: this file cannot be compiled in the build image (no Go toolchain); it is the binding a
// Quiver maintainer adds as pkg/gpu. Everything it calls is exercised through the same C ABI by
// the C++ host layer (quiver_b200/host) and the Python ctypes tests.
package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../quiver_b200/lib -lquivergpu -Wl,-rpath,${SRCDIR}/../../../quiver_b200/lib
#include <stdlib.h>
#include "quiver_gpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/TFMV/quiver/pkg/types"
	"github.com/TFMV/quiver/pkg/vectortypes"
)

// Metric mirrors qg_metric; it replaces the `%p` sniffing of db.go:322-334.
type Metric int

const (
	Cosine Metric = iota
	Euclidean
	DotProduct
	SquaredEuclidean
	Manhattan
)

// MetricOf maps the reference's DistanceType strings (vectortypes/types.go:19-25).
func MetricOf(t vectortypes.DistanceType) Metric {
	switch t {
	case vectortypes.Euclidean:
		return Euclidean
	case vectortypes.DotProduct:
		return DotProduct
	case vectortypes.Manhattan:
		return Manhattan
	default: // unknown => cosine, types.go:46-47
		return Cosine
	}
}

// call runs one qg_* call and, when it fails, reads the library's thread-local error text ON THE SAME OS
// THREAD: the Go scheduler may move a goroutine to another thread between two cgo calls, and the message
// ("k must be positive", "query dimension mismatch: expected %d, got %d", ...) would then come back empty or
// stale. The pair is therefore pinned with LockOSThread.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return fmt.Errorf("%s (qg_status %d)", C.GoString(C.qg_last_error()), int(rc))
	}
	return nil
}

// Index satisfies core.Index and core.BatchIndex (pkg/core/collection.go:78-96) and
// hybrid.Index (pkg/hybrid/types.go:196-214). String IDs live here; the device sees rows.
type Index struct {
	mu     sync.RWMutex // same discipline as exact.go:25: searches share, mutations exclude
	h      *C.qg_index
	dim    int
	ids    []string         // row -> id
	rows   map[string]int64 // id -> row (live rows only)
	metric Metric
}

// Options of New beyond the defaults. NoBF16Copy drops the bf16 copy of the corpus the library keeps for
// dim <= 512 (a third of the index's device memory): batches of 8 and more then run as tf32, smaller ones on
// the flat fp32 scan; results are identical either way.
type Options struct {
	NoBF16Copy  bool
	ReserveRows int64
}

func New(dim int, metric Metric, device int) (*Index, error) {
	return NewWithOptions(dim, metric, device, Options{})
}

func NewWithOptions(dim int, metric Metric, device int, opt Options) (*Index, error) {
	cfg := C.qg_config{device: C.int(device), reserve_rows: C.int64_t(opt.ReserveRows)}
	if opt.NoBF16Copy {
		cfg.flags = C.QG_FLAG_NO_BF16_COPY
	}
	var h *C.qg_index
	if err := call(func() C.int { return C.qg_index_create(&h, C.int(dim), C.int(metric), &cfg) }); err != nil {
		return nil, err
	}
	return &Index{h: h, dim: dim, rows: map[string]int64{}, metric: metric}, nil
}

func (x *Index) Close() { C.qg_index_destroy(x.h) }

// Insert — exact.go:38-58: dimension check, duplicate check, the vector is copied.
func (x *Index) Insert(id string, v vectortypes.F32) error {
	x.mu.Lock()
	defer x.mu.Unlock()
	if len(v) != x.dim {
		return fmt.Errorf("vector dimension mismatch: expected %d, got %d", x.dim, len(v))
	}
	if _, dup := x.rows[id]; dup {
		return fmt.Errorf("vector with ID %s already exists", id)
	}
	var first C.int64_t
	if err := call(func() C.int { return C.qg_index_upload(x.h, (*C.float)(unsafe.Pointer(&v[0])), 1, &first) }); err != nil {
		return err
	}
	x.ids = append(x.ids, id)
	x.rows[id] = int64(first)
	return nil
}

// InsertBatch — core.BatchIndex: one pinned-staged upload for the whole batch.
func (x *Index) InsertBatch(vs map[string]vectortypes.F32) error {
	x.mu.Lock()
	defer x.mu.Unlock()
	flat := make([]float32, 0, len(vs)*x.dim)
	ids := make([]string, 0, len(vs))
	for id, v := range vs {
		if len(v) != x.dim {
			return fmt.Errorf("vector dimension mismatch: expected %d, got %d", x.dim, len(v))
		}
		if _, dup := x.rows[id]; dup {
			return fmt.Errorf("vector with ID %s already exists", id)
		}
		flat = append(flat, v...)
		ids = append(ids, id)
	}
	if len(ids) == 0 {
		return nil
	}
	var first C.int64_t
	if err := call(func() C.int {
		return C.qg_index_upload(x.h, (*C.float)(unsafe.Pointer(&flat[0])), C.int64_t(len(ids)), &first)
	}); err != nil {
		return err
	}
	for i, id := range ids {
		x.ids = append(x.ids, id)
		x.rows[id] = int64(first) + int64(i)
	}
	return nil
}

// Delete — exact.go:61-70: deleting a missing id is a no-op. Rows become tombstones.
func (x *Index) Delete(id string) error {
	x.mu.Lock()
	defer x.mu.Unlock()
	row, ok := x.rows[id]
	if !ok {
		return nil
	}
	r := C.int64_t(row)
	if err := call(func() C.int { return C.qg_index_tombstone(x.h, &r, 1) }); err != nil {
		return err
	}
	delete(x.rows, id)
	// The reference's map frees a vector on Delete; here dead rows stay in HBM until they outnumber the
	// live ones, then one qg_index_compact pass squeezes them out (the write lock is already held).
	if dead := len(x.ids) - len(x.rows); dead >= compactMinDead && dead > len(x.rows) {
		return x.compactLocked()
	}
	return nil
}

const compactMinDead = 4096

// Compact drops tombstoned rows from the device arrays and renumbers the id <-> row tables.
func (x *Index) Compact() error {
	x.mu.Lock()
	defer x.mu.Unlock()
	return x.compactLocked()
}

func (x *Index) compactLocked() error {
	if len(x.ids) == len(x.rows) {
		return nil
	}
	oldToNew := make([]C.int64_t, len(x.ids))
	var n C.int64_t
	if err := call(func() C.int { return C.qg_index_compact(x.h, &oldToNew[0], &n) }); err != nil {
		return err
	}
	ids := make([]string, int(n))
	for r, j := range oldToNew {
		if j >= 0 {
			ids[j] = x.ids[r]
			x.rows[x.ids[r]] = int64(j)
		}
	}
	x.ids = ids
	return nil
}

// DeleteBatch — core.BatchIndex (collection.go:91-96): one tombstone launch for the whole batch.
// Missing ids are skipped like ExactIndex.Delete does (exact.go:61-70).
func (x *Index) DeleteBatch(ids []string) error {
	x.mu.Lock()
	defer x.mu.Unlock()
	rows := make([]C.int64_t, 0, len(ids))
	for _, id := range ids {
		if row, ok := x.rows[id]; ok {
			rows = append(rows, C.int64_t(row))
		}
	}
	if len(rows) == 0 {
		return nil
	}
	if err := call(func() C.int { return C.qg_index_tombstone(x.h, &rows[0], C.int64_t(len(rows))) }); err != nil {
		return err
	}
	for _, id := range ids {
		delete(x.rows, id)
	}
	if dead := len(x.ids) - len(x.rows); dead >= compactMinDead && dead > len(x.rows) {
		return x.compactLocked()
	}
	return nil
}

func (x *Index) Size() int { return int(C.qg_index_size(x.h)) }

// Search — exact.go:92-133.
func (x *Index) Search(q vectortypes.F32, k int) ([]types.BasicSearchResult, error) {
	res, err := x.BatchSearch([]vectortypes.F32{q}, k)
	if err != nil {
		return nil, err
	}
	return res[0], nil
}

// BatchSearch — hybrid_index.go:677-811 with ForceStrategy = exact: ONE call for all queries.
func (x *Index) BatchSearch(qs []vectortypes.F32, k int) ([][]types.BasicSearchResult, error) {
	x.mu.RLock()
	defer x.mu.RUnlock()
	n := len(qs)
	if n == 0 {
		return nil, nil
	}
	dim := len(qs[0])
	flat := make([]float32, 0, n*dim)
	for _, q := range qs {
		if len(q) != dim {
			return nil, errors.New("queries of one batch must share a dimension")
		}
		flat = append(flat, q...)
	}
	kk := k
	if kk < 1 {
		kk = 1
	}
	dist := make([]float32, n*kk)
	rows := make([]int64, n*kk)
	cnt := make([]int32, n)
	if err := call(func() C.int {
		return C.qg_search_batch(x.h, (*C.float)(unsafe.Pointer(&flat[0])), C.int(n), C.int(dim), C.int(k), nil, nil,
			(*C.float)(unsafe.Pointer(&dist[0])), nil, (*C.int64_t)(unsafe.Pointer(&rows[0])), (*C.int)(unsafe.Pointer(&cnt[0])))
	}); err != nil {
		return nil, err // "k must be positive", "query dimension mismatch: ..."
	}
	out := make([][]types.BasicSearchResult, n)
	for i := range out {
		out[i] = make([]types.BasicSearchResult, cnt[i])
		for j := 0; j < int(cnt[i]); j++ {
			out[i][j] = types.BasicSearchResult{ID: x.ids[rows[i*kk+j]], Distance: dist[i*kk+j]}
		}
	}
	return out, nil
}

// BatchDistance is the optional hook consulted by hnsw.searchLayer (hnsw.go:536-563) instead of
// one computeDistance call per neighbour.
func (x *Index) BatchDistance(q []float32, rows []uint32, out []float32) error {
	if len(rows) == 0 {
		return nil
	}
	return call(func() C.int {
		return C.qg_batch_distance(x.h, (*C.float)(unsafe.Pointer(&q[0])), C.int(len(q)),
			(*C.uint32_t)(unsafe.Pointer(&rows[0])), C.int(len(rows)), (*C.float)(unsafe.Pointer(&out[0])))
	})
}

// SearchExhaustive is ExactIndex.Search as the reference literally runs it — the exact distance of every live row,
// a full sort, the first k (exact.go:114-129) — on the device: no thresholds and no certificate. It is what
// UseExactSearch (search.go:12-31) maps to when a caller wants the slow path by name, and the GPU-side oracle of
// the parity tests.
func (x *Index) SearchExhaustive(q vectortypes.F32, k int) ([]types.BasicSearchResult, error) {
	x.mu.RLock()
	defer x.mu.RUnlock()
	dist := make([]float32, k)
	rows := make([]int64, k)
	var cnt C.int
	if err := call(func() C.int {
		return C.qg_search_exhaustive(x.h, (*C.float)(unsafe.Pointer(&q[0])), 1, C.int(len(q)), C.int(k), nil,
			(*C.float)(unsafe.Pointer(&dist[0])), (*C.int64_t)(unsafe.Pointer(&rows[0])), &cnt)
	}); err != nil {
		return nil, err
	}
	out := make([]types.BasicSearchResult, 0, int(cnt))
	for j := 0; j < int(cnt); j++ {
		out = append(out, types.BasicSearchResult{ID: x.ids[rows[j]], Distance: dist[j]})
	}
	return out, nil
}

// Graph is an HNSW graph resident on the device next to the index's rows: hnsw.Search (hnsw.go:602-713) runs as one
// kernel launch for a whole batch of queries (a warp per query, the reference's heaps, visit order and stop rules —
// step-identical to the host walk). Build constructs the graph on the device (batched inserts, the reference's
// neighbour selection and prune rule); Upload takes a graph the Go side already holds (HNSW.Nodes flattened).
type Graph struct {
	h   *C.qg_hnsw
	idx *Index
}

func (x *Index) BuildGraph(m, maxM0, efConstruction, maxLevel int, seed uint64) (*Graph, error) {
	x.mu.RLock()
	defer x.mu.RUnlock()
	var g *C.qg_hnsw
	if err := call(func() C.int {
		return C.qg_hnsw_build(x.h, C.int(m), C.int(maxM0), C.int(efConstruction), C.int(maxLevel), C.uint64_t(seed), 0, &g)
	}); err != nil {
		return nil, err
	}
	return &Graph{h: g, idx: x}, nil
}

// UploadGraph: level[i] = top layer of node i, adj0 = layer-0 lists (nodes x maxM0, 0xFFFFFFFF padded), the upper
// layers as one CSR (layout in quiver_gpu.h: qg_hnsw_upload; qg_hnsw_export writes the same arrays).
func (x *Index) UploadGraph(m, maxM0, entryPoint, currentLevel int, level []int32, adj0 []uint32, upperOff []int64,
	upperAdj []uint32) (*Graph, error) {
	var g *C.qg_hnsw
	var ua *C.uint32_t
	if len(upperAdj) > 0 {
		ua = (*C.uint32_t)(unsafe.Pointer(&upperAdj[0]))
	}
	if err := call(func() C.int {
		return C.qg_hnsw_upload(x.h, C.int64_t(len(level)), C.int(m), C.int(maxM0), C.int(entryPoint), C.int(currentLevel),
			(*C.int32_t)(unsafe.Pointer(&level[0])), (*C.uint32_t)(unsafe.Pointer(&adj0[0])),
			(*C.int64_t)(unsafe.Pointer(&upperOff[0])), ua, &g)
	}); err != nil {
		return nil, err
	}
	return &Graph{h: g, idx: x}, nil
}

// Search walks the graph for every query of the batch. count[i] < k: the graph walk under-filled, the caller runs
// the exact pass (hnsw.go:676-710 — Index.BatchSearch over those queries); count[i] < 0: the query's candidate heap
// outgrew the kernel's shared-memory slice, repeat it with the host walk (never seen at efSearch = 128).
func (g *Graph) Search(qs []float32, nq, k, efSearch int) (rows []uint32, dist []float32, count []int32, err error) {
	g.idx.mu.RLock()
	defer g.idx.mu.RUnlock()
	rows = make([]uint32, nq*k)
	dist = make([]float32, nq*k)
	count = make([]int32, nq)
	err = call(func() C.int {
		return C.qg_hnsw_search_batch(g.idx.h, g.h, (*C.float)(unsafe.Pointer(&qs[0])), C.int(nq), C.int(g.idx.dim), C.int(k),
			C.int(efSearch), (*C.uint32_t)(unsafe.Pointer(&rows[0])), (*C.float)(unsafe.Pointer(&dist[0])),
			(*C.int)(unsafe.Pointer(&count[0])), nil)
	})
	return
}

func (g *Graph) Close() { C.qg_hnsw_destroy(g.h) }

// Group drives every GPU of the box from this one process (qg_group_*: one worker thread, stream, index and
// NCCL communicator per device inside libquivergpu). It is what DB.BatchSearch (pkg/core/db.go:707-845) hands a
// batch to when the collection spans GPUs: row-sharded (every GPU scans its shard, one all-gather of the
// per-shard top-k keys, merge) or replicated with the batch split — the library picks by corpus size.
type Group struct {
	mu  sync.RWMutex
	h   *C.qg_group
	ids []string // global row -> id
	dim int
}

func NewGroup(devices []int, dim int, metric Metric) (*Group, error) {
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	var h *C.qg_group
	if err := call(func() C.int { return C.qg_group_create(&devs[0], C.int(len(devs)), C.int(dim), C.int(metric), nil, &h) }); err != nil {
		return nil, err
	}
	return &Group{h: h, dim: dim}, nil
}

// Load uploads the corpus once (rows in id order); layout 2 = QG_LAYOUT_AUTO.
func (g *Group) Load(ids []string, flat []float32) error {
	g.mu.Lock()
	defer g.mu.Unlock()
	if err := call(func() C.int {
		return C.qg_group_upload(g.h, (*C.float)(unsafe.Pointer(&flat[0])), C.int64_t(len(ids)), C.QG_LAYOUT_AUTO)
	}); err != nil {
		return err
	}
	g.ids = append([]string(nil), ids...)
	return nil
}

func (g *Group) BatchSearch(qs []vectortypes.F32, k int) ([][]types.BasicSearchResult, error) {
	g.mu.RLock()
	defer g.mu.RUnlock()
	n := len(qs)
	if n == 0 {
		return nil, nil
	}
	flat := make([]float32, 0, n*g.dim)
	for _, q := range qs {
		flat = append(flat, q...)
	}
	kk := k
	if kk < 1 {
		kk = 1
	}
	dist := make([]float32, n*kk)
	rows := make([]int64, n*kk)
	cnt := make([]int32, n)
	if err := call(func() C.int {
		return C.qg_group_search_batch(g.h, (*C.float)(unsafe.Pointer(&flat[0])), C.int(n), C.int(len(qs[0])), C.int(k),
			(*C.float)(unsafe.Pointer(&dist[0])), (*C.int64_t)(unsafe.Pointer(&rows[0])), (*C.int)(unsafe.Pointer(&cnt[0])))
	}); err != nil {
		return nil, err
	}
	out := make([][]types.BasicSearchResult, n)
	for i := range out {
		out[i] = make([]types.BasicSearchResult, cnt[i])
		for j := 0; j < int(cnt[i]); j++ {
			out[i][j] = types.BasicSearchResult{ID: g.ids[rows[i*kk+j]], Distance: dist[i*kk+j]}
		}
	}
	return out, nil
}

func (g *Group) Close() { C.qg_group_destroy(g.h) }
