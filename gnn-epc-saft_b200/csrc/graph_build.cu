// Device-side graph structure: destination CSR (+ implicit self-loops), source slot lists,
// graph pointers, and the packed-store collate.  Integer work, HBM/latency bound.
// Replaces PyG add_self_loops / scatter index handling / Batch.from_data_list
// (reference gnnepcsaft/train/models.py:118-121,133; train/train.py:74-79).
#include "common.cuh"

// SM count of the CURRENT device, cached per device ordinal (a process may drive several GPUs)
int gnn_sm_count() {
    static int cache[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cache[dev] = n;
        else cache[dev] = 148;
    }
    return cache[dev];
}

extern "C" const char* gnn_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case GNN_ERR_BAD_ARG: return "gnnepcsaft_b200: bad argument";
        case GNN_ERR_TOO_LARGE: return "gnnepcsaft_b200: problem too large for this kernel";
        case GNN_ERR_UNSUPPORTED: return "gnnepcsaft_b200: unsupported configuration";
        case GNN_ERR_WORKSPACE: return "gnnepcsaft_b200: workspace too small";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gnnepcsaft_b200: unknown error";
    }
}
extern "C" int gnn_abi_version(void) { return 1; }
extern "C" int gnn_device_info(int32_t* out_host) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return (int)e;
    out_host[0] = p.multiProcessorCount;
    out_host[1] = p.major;
    out_host[2] = p.minor;
    return 0;
}

// ------------------------------------------------------------------ exclusive scan (int32)
// 3-kernel scan: per-block scan of SCAN_ITEMS*SCAN_THREADS elements, scan of block totals by
// one block, add-back.  In place.
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(GNN_FULL_MASK, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < (int)(blockDim.x >> 5)) ? smem_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(GNN_FULL_MASK, winc, o);
            if (lane >= o) winc += n;
        }
        smem_warp[lane] = winc - w;       // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;
    }
    __syncthreads();
    block_total = smem_warp[32];
    int res = smem_warp[warp] + inc - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(int* data, int64_t n, int* block_sums) {
    __shared__ int sw[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        local += v[k];
    }
    int total;
    int off = block_exclusive_scan(local, sw, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = off;
        off += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(int* block_sums, int nblocks) {
    __shared__ int sw[33];
    int carry = 0;
    for (int start = 0; start < nblocks; start += SCAN_THREADS) {
        int i = start + threadIdx.x;
        int v = i < nblocks ? block_sums[i] : 0;
        int total;
        int off = block_exclusive_scan(v, sw, total);
        if (i < nblocks) block_sums[i] = off + carry;
        carry += total;
    }
}

__global__ void scan_add_kernel(int* data, int64_t n, const int* block_sums) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_sums[i / SCAN_TILE];
}

// exclusive scan of data[0..n) in place; block_sums needs ceil(n/SCAN_TILE) ints
static int exclusive_scan_i32(int* data, int64_t n, int* block_sums, cudaStream_t st) {
    if (n <= 0) return 0;
    int64_t nblocks = gnn_ceil_div(n, SCAN_TILE);
    if (nblocks > (1 << 30)) return GNN_ERR_TOO_LARGE;
    scan_tiles_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, st>>>(data, n, block_sums);
    if (nblocks > 1) {
        scan_block_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(block_sums, (int)nblocks);
        scan_add_kernel<<<(unsigned)gnn_ceil_div(n, 256), 256, 0, st>>>(data, n, block_sums);
    }
    return gnn_launch_status();
}

// ------------------------------------------------------------------ CSR build
__global__ void csr_init_counts_kernel(int* rowptr, int* srcptr, int* fill_in, int* fill_out,
                                       int64_t n, int init) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) {
        int v = i < n ? init : 0;
        rowptr[i] = v;
        srcptr[i] = v;
    }
    if (i < n) {
        fill_in[i] = 0;
        fill_out[i] = 0;
    }
}

__global__ void csr_count_kernel(const int64_t* __restrict__ edge_index, int64_t n, int64_t e,
                                 int* rowptr, int* srcptr, int* err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    int64_t s = edge_index[i], d = edge_index[e + i];
    if (s < 0 || s >= n || d < 0 || d >= n) {
        if (err) atomicOr(err, GNN_INPUT_ERR_EDGE_INDEX);
        return;
    }
    atomicAdd(&rowptr[d], 1);
    atomicAdd(&srcptr[s], 1);
}

__global__ void csr_fill_eid_kernel(const int64_t* __restrict__ edge_index, int64_t n, int64_t e,
                                    int self_loops, const int* __restrict__ rowptr, int* fill_in,
                                    int* col_eid) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = e + (self_loops ? n : 0);
    if (i >= total) return;
    int64_t d;
    if (i < e) {
        int64_t s = edge_index[i];
        d = edge_index[e + i];
        if (s < 0 || s >= n || d < 0 || d >= n) return;
    } else {
        d = i - e;
    }
    int pos = atomicAdd(&fill_in[d], 1);
    col_eid[rowptr[d] + pos] = (int)i;
}

// ascending insertion sort of every segment (segments are tiny: in-degree of an atom)
__global__ void sort_segments_kernel(const int* __restrict__ ptr, int64_t n, int* vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = ptr[i], e = ptr[i + 1];
    for (int a = b + 1; a < e; ++a) {
        int key = vals[a];
        int j = a - 1;
        while (j >= b && vals[j] > key) {
            vals[j + 1] = vals[j];
            --j;
        }
        vals[j + 1] = key;
    }
}

__global__ void csr_finalize_kernel(const int64_t* __restrict__ edge_index,
                                    const int64_t* __restrict__ edge_attr, int64_t n, int64_t e,
                                    int64_t total, const int* __restrict__ col_eid, int* col_src,
                                    uint8_t* col_code, const int* __restrict__ rowptr,
                                    const int* __restrict__ srcptr, int* fill_out, int* src_slot, int* err) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total || s >= rowptr[n]) return;   // fewer slots than edges only if edges were invalid
    int eid = col_eid[s];
    int src, code = 0;
    if (eid < e) {
        src = (int)edge_index[eid];
        int64_t a0 = edge_attr[3 * (int64_t)eid], a1 = edge_attr[3 * (int64_t)eid + 1],
                a2 = edge_attr[3 * (int64_t)eid + 2];
        if (a0 < 0 || a0 >= 5 || a1 < 0 || a1 >= 6 || a2 < 0 || a2 >= 2) {
            if (err) atomicOr(err, GNN_INPUT_ERR_BOND_FEATURE);
            a0 = a1 = a2 = 0;
        }
        code = (int)(a0 * 12 + a1 * 2 + a2);
    } else {
        src = eid - (int)e;
    }
    col_src[s] = src;
    col_code[s] = (uint8_t)code;
    int pos = atomicAdd(&fill_out[src], 1);
    src_slot[srcptr[src] + pos] = (int)s;
}

extern "C" int64_t gnn_build_csr_workspace_bytes(int64_t n, int64_t e) {
    (void)e;
    int64_t ints = 2 * (n + 1) + 2 * (gnn_ceil_div(n + 1, SCAN_TILE) + 1);
    return gnn_align_up(ints * 4, 256);
}

extern "C" int gnn_build_csr(const int64_t* edge_index, const int64_t* edge_attr, int64_t n,
                             int64_t e, int self_loops, int32_t* rowptr, int32_t* col_src,
                             uint8_t* col_code, int32_t* col_eid, int32_t* srcptr,
                             int32_t* src_slot, int32_t* err_flag, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    if (n < 0 || e < 0) return GNN_ERR_BAD_ARG;
    int64_t total = e + (self_loops ? n : 0);
    if (n >= (1ll << 31) - 1 || total >= (1ll << 31) - 1) return GNN_ERR_TOO_LARGE;
    if (workspace_bytes < gnn_build_csr_workspace_bytes(n, e)) return GNN_ERR_WORKSPACE;
    cudaStream_t st = gnn_stream(stream);
    int* ws = reinterpret_cast<int*>(workspace);
    int* fill_in = ws;
    int* fill_out = fill_in + (n + 1);
    int* bs_in = fill_out + (n + 1);
    int* bs_out = bs_in + gnn_ceil_div(n + 1, SCAN_TILE) + 1;
    const int T = 256;
    if (err_flag) cudaMemsetAsync(err_flag, 0, sizeof(int), st);
    csr_init_counts_kernel<<<(unsigned)gnn_ceil_div(n + 1, T), T, 0, st>>>(rowptr, srcptr, fill_in, fill_out, n,
                                                                             self_loops ? 1 : 0);
    if (e > 0)
        csr_count_kernel<<<(unsigned)gnn_ceil_div(e, T), T, 0, st>>>(edge_index, n, e, rowptr, srcptr, err_flag);
    int rc = exclusive_scan_i32(rowptr, n + 1, bs_in, st);
    if (rc) return rc;
    rc = exclusive_scan_i32(srcptr, n + 1, bs_out, st);
    if (rc) return rc;
    if (total > 0) {
        csr_fill_eid_kernel<<<(unsigned)gnn_ceil_div(total, T), T, 0, st>>>(edge_index, n, e, self_loops, rowptr,
                                                                             fill_in, col_eid);
        if (n > 0) sort_segments_kernel<<<(unsigned)gnn_ceil_div(n, T), T, 0, st>>>(rowptr, n, col_eid);
        csr_finalize_kernel<<<(unsigned)gnn_ceil_div(total, T), T, 0, st>>>(
            edge_index, edge_attr, n, e, total, col_eid, col_src, col_code, rowptr, srcptr, fill_out, src_slot,
            err_flag);
        if (n > 0) sort_segments_kernel<<<(unsigned)gnn_ceil_div(n, T), T, 0, st>>>(srcptr, n, src_slot);
    }
    return gnn_launch_status();
}

// ------------------------------------------------------------------ graph_ptr from batch
__global__ void graph_count_kernel(const int64_t* __restrict__ batch, int64_t n, int64_t b,
                                   int* node_graph, int* graph_ptr, int* err_flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t g = batch[i];
    // the reference's scatter-based global_add_pool raises on an out-of-range id and accepts any order; this path
    // needs the sorted vector PyG's Batch produces: both violations are reported through err_flag
    if (err_flag != nullptr) {
        if (g < 0 || g >= b) atomicOr(err_flag, GNN_INPUT_ERR_BATCH_RANGE);
        else if (i > 0 && batch[i - 1] > g) atomicOr(err_flag, GNN_INPUT_ERR_BATCH_UNSORTED);
    }
    if (g < 0) g = 0;
    if (g >= b) g = b - 1;
    node_graph[i] = (int)g;
    // batch is sorted (PyG Batch): graph g starts where batch changes
    if (i == 0 || batch[i - 1] != batch[i]) {
        // every graph in (prev, g] that has no nodes also starts here
        int64_t prev = i == 0 ? -1 : batch[i - 1];
        if (prev < -1) prev = -1;
        for (int64_t k = prev + 1; k <= g; ++k) graph_ptr[k] = (int)i;
    }
    if (i == n - 1)
        for (int64_t k = g + 1; k <= b; ++k) graph_ptr[k] = (int)n;
}
__global__ void graph_ptr_empty_kernel(int* graph_ptr, int64_t b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= b) graph_ptr[i] = 0;
}

extern "C" int gnn_graph_ptr_from_batch(const int64_t* batch, int64_t n, int64_t b, int32_t* node_graph,
                                        int32_t* graph_ptr, int32_t* err_flag, void* stream) {
    if (n < 0 || b < 1 || n >= (1ll << 31)) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    if (n == 0) {
        graph_ptr_empty_kernel<<<(unsigned)gnn_ceil_div(b + 1, 256), 256, 0, st>>>(graph_ptr, b);
        return gnn_launch_status();
    }
    graph_count_kernel<<<(unsigned)gnn_ceil_div(n, 256), 256, 0, st>>>(batch, n, b, node_graph, graph_ptr, err_flag);
    return gnn_launch_status();
}

// single molecule (data.batch is None): graph_ptr = [0, n]
__global__ void graph_ptr_single_kernel(int* graph_ptr, int n) {
    graph_ptr[0] = 0;
    graph_ptr[1] = n;
}
extern "C" int gnn_graph_ptr_single(int64_t n, int32_t* graph_ptr, void* stream) {
    if (n < 0 || n >= (1ll << 31)) return GNN_ERR_BAD_ARG;
    graph_ptr_single_kernel<<<1, 1, 0, gnn_stream(stream)>>>(graph_ptr, (int)n);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ collate from a packed store
// Offsets: counts -> exclusive scans (int64 outputs; the scan itself runs on int32 counts).
__global__ void collate_counts_kernel(const int64_t* __restrict__ ids, int64_t nb,
                                      const int64_t* __restrict__ node_slices,
                                      const int64_t* __restrict__ edge_slices, int* ncnt, int* ecnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nb) return;
    if (i == nb) {
        ncnt[i] = 0;
        ecnt[i] = 0;
        return;
    }
    int64_t g = ids[i];
    ncnt[i] = (int)(node_slices[g + 1] - node_slices[g]);
    ecnt[i] = (int)(edge_slices[g + 1] - edge_slices[g]);
}
__global__ void widen_kernel(const int* __restrict__ a, const int* __restrict__ b, int64_t n, int64_t* oa,
                             int64_t* ob) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        oa[i] = a[i];
        ob[i] = b[i];
    }
}

extern "C" int64_t gnn_collate_workspace_bytes(int64_t nb) {
    int64_t ints = 2 * (nb + 1) + 2 * (gnn_ceil_div(nb + 1, SCAN_TILE) + 1);
    return gnn_align_up(ints * 4, 256);
}
extern "C" int gnn_collate_offsets(const int64_t* ids, int64_t nb, const int64_t* node_slices,
                                   const int64_t* edge_slices, int64_t* out_ptr, int64_t* out_eptr,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    if (nb < 0) return GNN_ERR_BAD_ARG;
    if (workspace_bytes < gnn_collate_workspace_bytes(nb)) return GNN_ERR_WORKSPACE;
    cudaStream_t st = gnn_stream(stream);
    int* ncnt = reinterpret_cast<int*>(workspace);
    int* ecnt = ncnt + (nb + 1);
    int* bs0 = ecnt + (nb + 1);
    int* bs1 = bs0 + gnn_ceil_div(nb + 1, SCAN_TILE) + 1;
    collate_counts_kernel<<<(unsigned)gnn_ceil_div(nb + 1, 256), 256, 0, st>>>(ids, nb, node_slices, edge_slices,
                                                                                ncnt, ecnt);
    int rc = exclusive_scan_i32(ncnt, nb + 1, bs0, st);
    if (rc) return rc;
    rc = exclusive_scan_i32(ecnt, nb + 1, bs1, st);
    if (rc) return rc;
    widen_kernel<<<(unsigned)gnn_ceil_div(nb + 1, 256), 256, 0, st>>>(ncnt, ecnt, nb + 1, out_ptr, out_eptr);
    return gnn_launch_status();
}

// one warp per selected graph: copies its node rows / edge rows, adds the node offset to
// edge_index and writes the batch vector.
__global__ void collate_gather_kernel(const int64_t* __restrict__ ids, int64_t nb,
                                      const int64_t* __restrict__ node_slices,
                                      const int64_t* __restrict__ edge_slices,
                                      const int64_t* __restrict__ sx, const int64_t* __restrict__ sei,
                                      int64_t store_e, const int64_t* __restrict__ sea,
                                      const int64_t* __restrict__ optr, const int64_t* __restrict__ oeptr,
                                      int64_t tot_e, int64_t* x, int64_t* ei, int64_t* ea, int64_t* batch) {
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = w; b < nb; b += nwarps) {
        int64_t g = ids[b];
        int64_t n0 = node_slices[g], n1 = node_slices[g + 1];
        int64_t e0 = edge_slices[g], e1 = edge_slices[g + 1];
        int64_t on = optr[b], oe = oeptr[b];
        int64_t nn = n1 - n0, ne = e1 - e0;
        for (int64_t k = lane; k < nn * GNN_NUM_ATOM_FEATS; k += 32) x[on * GNN_NUM_ATOM_FEATS + k] = sx[n0 * GNN_NUM_ATOM_FEATS + k];
        for (int64_t k = lane; k < nn; k += 32) batch[on + k] = b;
        for (int64_t k = lane; k < ne * GNN_NUM_BOND_FEATS; k += 32) ea[oe * GNN_NUM_BOND_FEATS + k] = sea[e0 * GNN_NUM_BOND_FEATS + k];
        for (int64_t k = lane; k < ne; k += 32) {
            ei[oe + k] = sei[e0 + k] + on;
            ei[tot_e + oe + k] = sei[store_e + e0 + k] + on;
        }
    }
}

extern "C" int gnn_collate_gather(const int64_t* ids, int64_t nb, const int64_t* node_slices,
                                  const int64_t* edge_slices, const int64_t* store_x,
                                  const int64_t* store_edge_index, int64_t store_num_edges,
                                  const int64_t* store_edge_attr, const int64_t* out_ptr,
                                  const int64_t* out_eptr, int64_t total_nodes, int64_t total_edges,
                                  int64_t* x, int64_t* edge_index, int64_t* edge_attr, int64_t* batch,
                                  void* stream) {
    (void)total_nodes;
    if (nb <= 0) return nb == 0 ? 0 : GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    int64_t blocks = gnn_ceil_div(nb, 8);
    int64_t cap = (int64_t)gnn_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    collate_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(ids, nb, node_slices, edge_slices, store_x,
                                                            store_edge_index, store_num_edges, store_edge_attr,
                                                            out_ptr, out_eptr, total_edges, x, edge_index,
                                                            edge_attr, batch);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ compact collate: store CSR -> batch CSR
// The device store keeps ONE destination CSR over the whole dataset (built once by gnn_build_csr on the concatenated
// graphs: no edge crosses a graph, so a graph's nodes / CSR slots are contiguous ranges).  A batch is then a pure
// segment gather with re-based indices - no sort, no scan over edges - that emits what the model kernels consume
// directly: uint8 atom features, int32 CSR, uint8 bond codes, node -> graph map, graph_ptr.  Replaces the PyG
// DataLoader / Collater (reference train.py:74-79) AND add_self_loops + the scatter index handling (models.py:118-121)
// for batches drawn from the store; the int64 PyG-layout path (gnn_collate_gather + gnn_build_csr) stays for batches
// that arrive as PyG tensors.
// Padding: the outputs are sized (pad_nodes, pad_slots) >= (N, E') and hold cap >= B graphs; graphs B .. cap-1 are
// empty and the node / slot surplus forms one extra graph (index cap) of isolated nodes that carry the surplus CSR
// slots as self-edges, so every batch of a shape bucket runs the same launches (CUDA-graph replay).
// valid = {N, B, E'} tells the BatchNorm / loss kernels which rows are real.
__global__ void __launch_bounds__(1024)
compact_offsets_kernel(const int64_t* __restrict__ ids, int64_t nb, int64_t num_store_graphs,
                       const int* __restrict__ node_off, const int* __restrict__ slot_off, int* __restrict__ nptr,
                       int* __restrict__ sptr, int* err) {
    __shared__ int wn[32], ws[32];
    __shared__ int carry_n, carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        carry_n = 0;
        carry_s = 0;
        nptr[0] = 0;
        sptr[0] = 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int cn = 0, cs = 0;
        if (i < nb) {
            int64_t g = ids[i];
            if (g < 0 || g >= num_store_graphs) {
                if (err) atomicOr(err, GNN_INPUT_ERR_BATCH_RANGE);
                g = 0;
            }
            cn = node_off[g + 1] - node_off[g];
            cs = slot_off[g + 1] - slot_off[g];
        }
        int sn = cn, ss = cs;                       // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tn = __shfl_up_sync(GNN_FULL_MASK, sn, o), ts = __shfl_up_sync(GNN_FULL_MASK, ss, o);
            if (lane >= o) {
                sn += tn;
                ss += ts;
            }
        }
        if (lane == 31) {
            wn[warp] = sn;
            ws[warp] = ss;
        }
        __syncthreads();
        if (warp == 0) {
            int a = wn[lane], b = ws[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(GNN_FULL_MASK, a, o), tb = __shfl_up_sync(GNN_FULL_MASK, b, o);
                if (lane >= o) {
                    a += ta;
                    b += tb;
                }
            }
            wn[lane] = a;
            ws[lane] = b;
        }
        __syncthreads();
        const int pn = carry_n + (warp > 0 ? wn[warp - 1] : 0), ps = carry_s + (warp > 0 ? ws[warp - 1] : 0);
        if (i < nb) {
            nptr[i + 1] = pn + sn;
            sptr[i + 1] = ps + ss;
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry_n = pn + sn;
            carry_s = ps + ss;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
compact_gather_kernel(const int64_t* __restrict__ ids, int64_t nb, int64_t num_store_graphs,
                      const int* __restrict__ node_off, const int* __restrict__ slot_off,
                      const uint8_t* __restrict__ sx, const int* __restrict__ s_rowptr, const int* __restrict__ s_col_src,
                      const uint8_t* __restrict__ s_col_code, const float* __restrict__ s_para, int num_para,
                      const int* __restrict__ nptr, const int* __restrict__ sptr, int64_t pad_nodes, int64_t pad_slots,
                      uint8_t* __restrict__ x, int* __restrict__ rowptr, int* __restrict__ col_src,
                      uint8_t* __restrict__ col_code, int* __restrict__ node_graph, int* __restrict__ graph_ptr,
                      float* __restrict__ para, int* err) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    const bool fits = nptr[nb] <= pad_nodes && sptr[nb] <= pad_slots;
    if (!fits) {
        if (err && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, GNN_INPUT_ERR_EDGE_INDEX);
        return;
    }
    for (int64_t i = warp0; i < nb; i += nwarps) {
        int64_t g = ids[i];
        if (g < 0 || g >= num_store_graphs) g = 0;
        const int n0 = node_off[g], cn = node_off[g + 1] - n0, s0 = slot_off[g], cs = slot_off[g + 1] - s0;
        const int on = nptr[i], os = sptr[i];
        for (int k = lane; k < cn * GNN_NUM_ATOM_FEATS; k += 32)
            x[(int64_t)on * GNN_NUM_ATOM_FEATS + k] = __ldg(sx + (int64_t)n0 * GNN_NUM_ATOM_FEATS + k);
        for (int k = lane; k < cn; k += 32) {
            rowptr[on + k] = __ldg(s_rowptr + n0 + k) - s0 + os;
            node_graph[on + k] = (int)i;
        }
        for (int k = lane; k < cs; k += 32) {
            col_src[os + k] = __ldg(s_col_src + s0 + k) - n0 + on;
            col_code[os + k] = __ldg(s_col_code + s0 + k);
        }
        if (lane == 0) graph_ptr[i] = on;
        if (para != nullptr)
            for (int k = lane; k < num_para; k += 32) para[i * num_para + k] = __ldg(s_para + g * num_para + k);
    }
}

__global__ void __launch_bounds__(256)
compact_pad_kernel(int64_t nb, int64_t cap, const int* __restrict__ nptr, const int* __restrict__ sptr, int64_t pad_nodes,
                   int64_t pad_slots, int own_surplus, int num_para, uint8_t* __restrict__ x, int* __restrict__ rowptr,
                   int* __restrict__ col_src, uint8_t* __restrict__ col_code, int* __restrict__ node_graph,
                   int* __restrict__ graph_ptr, float* __restrict__ para, int* __restrict__ valid, int* err) {
    const int n = nptr[nb], e = sptr[nb];
    if (n > pad_nodes || e > pad_slots) return;          // reported by the gather kernel
    const int nd = (int)pad_nodes - n, d = (int)pad_slots - e;      // padding nodes / surplus slots
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid == 0) {
        rowptr[pad_nodes] = own_surplus ? (int)pad_slots : e;
        graph_ptr[cap + 1] = (int)pad_nodes;
        valid[0] = n;
        valid[1] = (int)nb;
        valid[2] = e;
        valid[3] = 0;
        if (own_surplus && nd == 0 && d != 0 && err) atomicOr(err, GNN_INPUT_ERR_EDGE_INDEX);   // surplus slots need a padding node
    }
    // graphs nb .. cap-1 are empty (a batch with fewer graphs than the bucket holds), graph `cap` is the padding graph
    for (int64_t g = nb + tid; g <= cap; g += nth) graph_ptr[g] = n;
    if (para != nullptr)
        for (int64_t k = nb * num_para + tid; k < (cap + 1) * num_para; k += nth) para[k] = 1.0f;   // unused targets
    if (nd == 0) return;
    // own_surplus: the surplus slots are spread over the padding nodes as self-edges (models with per-slot tensors:
    // every slot of the padded arrays then belongs to a node); otherwise the padding nodes have no in-edges and the
    // surplus slots are simply unused capacity
    const int base = own_surplus ? d / nd : 0, rem = own_surplus ? d % nd : 0;
    for (int k = tid; k < nd; k += nth) {
        const int node = n + k, start = e + k * base + min(k, rem), cnt = base + (k < rem ? 1 : 0);
        for (int j = 0; j < GNN_NUM_ATOM_FEATS; ++j) x[(int64_t)node * GNN_NUM_ATOM_FEATS + j] = 0;
        rowptr[node] = start;
        node_graph[node] = (int)cap;
        for (int j = 0; j < cnt; ++j) {
            col_src[start + j] = node;
            col_code[start + j] = 0;
        }
    }
}

extern "C" int64_t gnn_compact_collate_workspace_bytes(int64_t num_ids) {
    return gnn_align_up(2 * (num_ids + 1) * (int64_t)sizeof(int), 256);
}
extern "C" int gnn_compact_collate(const int64_t* ids, int64_t nb, int64_t cap, int64_t num_store_graphs, const int32_t* node_off,
                                   const int32_t* slot_off, const uint8_t* store_x, const int32_t* store_rowptr,
                                   const int32_t* store_col_src, const uint8_t* store_col_code, const float* store_para,
                                   int num_para, int64_t pad_nodes, int64_t pad_slots, int own_surplus, uint8_t* x,
                                   int32_t* rowptr,
                                   int32_t* col_src, uint8_t* col_code, int32_t* node_graph, int32_t* graph_ptr,
                                   float* para, int32_t* valid, int32_t* err_flag, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
    if (nb < 1 || cap < nb || num_store_graphs < 1 || pad_nodes < 0 || pad_slots < 0 || num_para < 0) return GNN_ERR_BAD_ARG;
    if (pad_nodes >= (1ll << 31) || pad_slots >= (1ll << 31)) return GNN_ERR_TOO_LARGE;
    if (workspace_bytes < gnn_compact_collate_workspace_bytes(nb)) return GNN_ERR_WORKSPACE;
    cudaStream_t st = gnn_stream(stream);
    int* nptr = reinterpret_cast<int*>(workspace);
    int* sptr = nptr + nb + 1;
    if (err_flag) cudaMemsetAsync(err_flag, 0, sizeof(int), st);
    compact_offsets_kernel<<<1, 1024, 0, st>>>(ids, nb, num_store_graphs, node_off, slot_off, nptr, sptr, err_flag);
    int64_t blocks = gnn_ceil_div(nb, 8);
    const int64_t max_blocks = (int64_t)gnn_sm_count() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    compact_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(ids, nb, num_store_graphs, node_off, slot_off, store_x,
                                                            store_rowptr, store_col_src, store_col_code, store_para,
                                                            num_para, nptr, sptr, pad_nodes, pad_slots, x, rowptr, col_src,
                                                            col_code, node_graph, graph_ptr, para, err_flag);
    compact_pad_kernel<<<32, 256, 0, st>>>(nb, cap, nptr, sptr, pad_nodes, pad_slots, own_surplus, num_para, x, rowptr, col_src, col_code,
                                           node_graph, graph_ptr, para, valid, err_flag);
    return gnn_launch_status();
}

// int64 [N, 9] OGB atom features -> uint8 (store build); values outside [0, 255] are reported
__global__ void x_to_u8_kernel(const int64_t* __restrict__ x, int64_t total, uint8_t* __restrict__ out, int* err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int64_t v = x[i];
        if ((v < 0 || v > 255) && err) atomicOr(err, GNN_INPUT_ERR_ATOM_FEATURE);
        out[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}
extern "C" int gnn_atom_features_to_u8(const int64_t* x, int64_t num_nodes, uint8_t* out, int32_t* err_flag, void* stream) {
    if (num_nodes < 0) return GNN_ERR_BAD_ARG;
    if (num_nodes == 0) return 0;
    const int64_t total = num_nodes * GNN_NUM_ATOM_FEATS;
    int64_t blocks = gnn_ceil_div(total, 256);
    int64_t cap = (int64_t)gnn_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    x_to_u8_kernel<<<(unsigned)blocks, 256, 0, gnn_stream(stream)>>>(x, total, out, err_flag);
    return gnn_launch_status();
}
