// OGB AtomEncoder and the bond-code table (BondEncoder -> PNAConv.edge_encoder -> bond third of
// pre_nns[t][0]).  Replace ogb.graphproppred.mol_encoder.{AtomEncoder,BondEncoder}
// (reference gnnepcsaft/train/models.py:65-66,122-123) and the edge_encoder / edge part of
// PNAConv.message (PyG nn/conv/pna_conv.py).
#include "common.cuh"

// ------------------------------------------------------------------ AtomEncoder forward
// IDX = int64_t (the reference's PyG batches) or uint8_t (compact batches of the device store: every OGB atom feature
// is < 119)
template <class IDX>
__global__ void atom_embed_fwd_kernel(const IDX* __restrict__ x, int64_t n, const float* __restrict__ tables,
                                      const int* __restrict__ row_off, int total_rows, int H,
                                      float* __restrict__ out, int* err_flag) {
    const int H4 = H >> 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * H4) return;
    int64_t i = idx / H4;
    int c = (int)(idx - i * H4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < GNN_NUM_ATOM_FEATS; ++k) {
        int lo = __ldg(row_off + k);
        int hi = (k + 1 < GNN_NUM_ATOM_FEATS) ? __ldg(row_off + k + 1) : total_rows;
        int64_t v = (int64_t)__ldg(x + i * GNN_NUM_ATOM_FEATS + k);
        if ((v < 0 || v >= hi - lo) && err_flag != nullptr && c == 0) atomicOr(err_flag, GNN_INPUT_ERR_ATOM_FEATURE);   // torch would raise IndexError
        int row = lo + (int)(v < 0 ? 0 : (v >= hi - lo ? hi - lo - 1 : v));
        float4 e = ldg4(tables + (int64_t)row * H + c);
        acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
    }
    st4(out + i * H + c, acc);
}

template <class IDX>
static int atom_embed_fwd_launch(const IDX* x, int64_t n, const float* tables, const int32_t* row_off, int32_t total_rows,
                                 int hidden, float* out, int32_t* err_flag, void* stream) {
    if (n < 0 || hidden < 4 || (hidden & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    int64_t work = n * (hidden >> 2);
    atom_embed_fwd_kernel<IDX><<<(unsigned)gnn_ceil_div(work, 256), 256, 0, gnn_stream(stream)>>>(
        x, n, tables, row_off, total_rows, hidden, out, err_flag);
    return gnn_launch_status();
}
extern "C" int gnn_atom_embed_fwd(const int64_t* x, int64_t n, const float* tables, const int32_t* row_off,
                                  int32_t total_rows, int hidden, float* out, int32_t* err_flag, void* stream) {
    return atom_embed_fwd_launch(x, n, tables, row_off, total_rows, hidden, out, err_flag, stream);
}
extern "C" int gnn_atom_embed_u8_fwd(const uint8_t* x, int64_t n, const float* tables, const int32_t* row_off,
                                     int32_t total_rows, int hidden, float* out, int32_t* err_flag, void* stream) {
    return atom_embed_fwd_launch(x, n, tables, row_off, total_rows, hidden, out, err_flag, stream);
}

// ------------------------------------------------------------------ AtomEncoder backward
// CTA-private copy of the (small) table gradient in shared memory, flushed once per CTA.
template <bool SMEM, class IDX>
__global__ void __launch_bounds__(256)
atom_embed_bwd_kernel(const IDX* __restrict__ x, int64_t n, const float* __restrict__ dout,
                      const int* __restrict__ row_off, int total_rows, int H, float* dtables) {
    extern __shared__ float s_tab[];
    if (SMEM) {
        for (int k = threadIdx.x; k < total_rows * H; k += blockDim.x) s_tab[k] = 0.f;
        __syncthreads();
    }
    float* acc = SMEM ? s_tab : dtables;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t i = warp0; i < n; i += nwarps) {
        int row = 0;
        if (lane < GNN_NUM_ATOM_FEATS) {
            int lo = __ldg(row_off + lane);
            int hi = (lane + 1 < GNN_NUM_ATOM_FEATS) ? __ldg(row_off + lane + 1) : total_rows;
            int64_t v = (int64_t)__ldg(x + i * GNN_NUM_ATOM_FEATS + lane);
            row = lo + (int)(v < 0 ? 0 : (v >= hi - lo ? hi - lo - 1 : v));
        }
        int rows[GNN_NUM_ATOM_FEATS];
#pragma unroll
        for (int k = 0; k < GNN_NUM_ATOM_FEATS; ++k) rows[k] = __shfl_sync(GNN_FULL_MASK, row, k);
        for (int c = lane; c < H; c += 32) {
            float g = __ldg(dout + i * H + c);
#pragma unroll
            for (int k = 0; k < GNN_NUM_ATOM_FEATS; ++k) atomicAdd(acc + rows[k] * H + c, g);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int k = threadIdx.x; k < total_rows * H; k += blockDim.x) {
            float v = s_tab[k];
            if (v != 0.f) atomicAdd(dtables + k, v);
        }
    }
}

// H = 64 (the default model): no atomics while accumulating.  sm_100 has no native fp32 shared-memory atomic add (it is
// a CAS loop) and almost every atom hits the same few rows (carbon, degree 3, ...).  Here ONE warp owns a node and one
// of four private copies of the table gradient (4 x 44 KB): the lower half-warp adds the node's 64 columns (a float4
// per lane) to the rows of features 0-4, the upper half to those of features 5-8 with plain loads and stores - the
// nine rows of a node lie in nine disjoint row ranges, so the read-modify-writes are independent; the copies are
// summed and flushed by 128-bit reductions.  Alone on the GPU: 37-47 us for 102 k nodes and 13 us for 12.8 k (atomic
// kernel 62 / 42 us); without the ring (registers, one trip ahead) it stayed at 62 us whatever the shared-memory
// instruction count was.
template <class IDX>
__global__ void __launch_bounds__(128, 1)
atom_embed_bwd_h64_kernel(const IDX* __restrict__ x, int64_t n, const float* __restrict__ dout,
                          const int* __restrict__ row_off, int total_rows, float* dtables) {
    extern __shared__ float4 s_tab4[];
    const int tab4 = total_rows * 16;                                       // float4 per copy (64 columns per row)
    for (int k = threadIdx.x; k < 4 * tab4; k += 128) s_tab4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, c4 = lane & 15;
    float4* acc = s_tab4 + warp * tab4 + c4;
    int lo = 0, span = 1;
    if (lane < GNN_NUM_ATOM_FEATS) {
        lo = __ldg(row_off + lane);
        span = ((lane + 1 < GNN_NUM_ATOM_FEATS) ? __ldg(row_off + lane + 1) : total_rows) - lo;
    }
    // dout rows travel global -> shared by cp.async into a per-warp ring of 4 trips x 8 rows (32 KB for the CTA), three
    // trips ahead: with four warps per SM the loop is bound by the latency of these loads (without them the kernel
    // took 29 instead of 62 us), so ~24 rows per warp have to be in flight, more than registers hold.
    constexpr int UNR = 8, DEPTH = 3;
    const int64_t stride = (int64_t)gridDim.x * 4;
    float4* ring = s_tab4 + 4 * tab4 + warp * ((DEPTH + 1) * UNR * 16);
    auto issue = [&](int64_t base, int slot0, int* rw) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = base + u * stride;
            rw[u] = 0;
            if (i < n && lane < GNN_NUM_ATOM_FEATS) {
                const int64_t v = (int64_t)__ldg(x + i * GNN_NUM_ATOM_FEATS + lane);
                rw[u] = (lo + (int)(v < 0 ? 0 : (v >= span ? span - 1 : v))) * 16;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; u += 2) {                     // one row per half-warp and instruction
            const int64_t i = base + (u + half) * stride;
            if (i < n) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (slot0 + u + half) * 16 + c4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(dout + i * 64 + c4 * 4) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int row0[UNR], row1[UNR], row2[UNR], row3[UNR];
    int64_t base = (int64_t)blockIdx.x * 4 + warp;
    const int64_t trip = stride * UNR;
    issue(base, 0, row0);
    issue(base + trip, UNR, row1);
    issue(base + 2 * trip, 2 * UNR, row2);
    for (int t = 0; base < n; base += trip, ++t) {
        __syncwarp();                                          // the slots refilled below were read in the last trip
        issue(base + DEPTH * trip, ((t + DEPTH) & 3) * UNR, row3);
        asm volatile("cp.async.wait_group 3;" ::: "memory");
        __syncwarp();
        const float4* gs = ring + (t & 3) * UNR * 16 + c4;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (base + u * stride < n) {                       // warp-uniform
                int r[5];
                float4 v[5];
                const float4 g = gs[u * 16];
#pragma unroll
                for (int j = 0; j < 5; ++j) r[j] = __shfl_sync(GNN_FULL_MASK, row0[u], min(half * 5 + j, GNN_NUM_ATOM_FEATS - 1));
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (j < 4 || half == 0) v[j] = acc[r[j]];
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (j < 4 || half == 0) acc[r[j]] = make_float4(v[j].x + g.x, v[j].y + g.y, v[j].z + g.z, v[j].w + g.w);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            row0[u] = row1[u];
            row1[u] = row2[u];
            row2[u] = row3[u];
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int k = threadIdx.x; k < tab4; k += 128) {
        float4 t = s_tab4[k];
#pragma unroll
        for (int c = 1; c < 4; ++c) {
            const float4 o = s_tab4[c * tab4 + k];
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        if (t.x != 0.f || t.y != 0.f || t.z != 0.f || t.w != 0.f) atomicAdd(reinterpret_cast<float4*>(dtables) + k, t);
    }
}

template <class IDX>
static int atom_embed_bwd_launch(const IDX* x, int64_t n, const float* dout, const int32_t* row_off, int32_t total_rows,
                                 int hidden, float* dtables, void* stream) {
    if (n < 0 || hidden < 1) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    cudaStream_t st = gnn_stream(stream);
    size_t smem = (size_t)total_rows * hidden * sizeof(float);
    if (hidden == 64 && smem * 4 + 32 * 1024 <= 227 * 1024 && (reinterpret_cast<uintptr_t>(dtables) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dout) & 15) == 0) {
        {
            auto kern = atom_embed_bwd_h64_kernel<IDX>;
            const size_t bytes = smem * 4 + 4 * 4 * 8 * 256;           // four table copies + the cp.async rings
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e != cudaSuccess) return (int)e;
            int64_t blocks = gnn_ceil_div(n, 4 * 8);
            if (blocks > gnn_sm_count()) blocks = gnn_sm_count();
            kern<<<(unsigned)blocks, 128, bytes, st>>>(x, n, dout, row_off, total_rows, dtables);
            return gnn_launch_status();
        }
    }
    int64_t blocks = gnn_ceil_div(n, 8 * 16);
    if (smem <= 200 * 1024) {
        auto kern = atom_embed_bwd_kernel<true, IDX>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        int per_sm = (int)((220 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        int64_t cap = (int64_t)gnn_sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        kern<<<(unsigned)blocks, 256, smem, st>>>(x, n, dout, row_off, total_rows, hidden, dtables);
    } else {
        int64_t cap = (int64_t)gnn_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        atom_embed_bwd_kernel<false, IDX><<<(unsigned)blocks, 256, 0, st>>>(x, n, dout, row_off, total_rows, hidden, dtables);
    }
    return gnn_launch_status();
}
extern "C" int gnn_atom_embed_bwd(const int64_t* x, int64_t n, const float* dout, const int32_t* row_off,
                                  int32_t total_rows, int hidden, float* dtables, void* stream) {
    return atom_embed_bwd_launch(x, n, dout, row_off, total_rows, hidden, dtables, stream);
}
extern "C" int gnn_atom_embed_u8_bwd(const uint8_t* x, int64_t n, const float* dout, const int32_t* row_off,
                                     int32_t total_rows, int hidden, float* dtables, void* stream) {
    return atom_embed_bwd_launch(x, n, dout, row_off, total_rows, hidden, dtables, stream);
}

// ------------------------------------------------------------------ bond-code table
// bond tables are the three BondEncoder embeddings stacked: rows [0,5) type, [5,11) stereo,
// [11,13) conjugated.  code = c0*12 + c1*2 + c2.
__device__ __forceinline__ void bond_rows(int code, int& r0, int& r1, int& r2) {
    r0 = code / 12;
    r1 = 5 + (code % 12) / 2;
    r2 = 11 + (code & 1);
}

// grid (60, L); one CTA per (code, layer)
__global__ void __launch_bounds__(128)
bond_table_fwd_kernel(const float* __restrict__ bond_tables, const float* __restrict__ wee,
                      const float* __restrict__ bee, const float* __restrict__ wpre, int64_t wpre_tower_stride,
                      int64_t layer_stride, int H, int towers, float* __restrict__ enc, float* __restrict__ tab) {
    extern __shared__ float sm[];
    float* s_emb = sm;         // [H]
    float* s_enc = sm + H;     // [F] (F == H, divide_input=False)
    const int code = blockIdx.x, layer = blockIdx.y;
    const int F = H;
    wee += layer * layer_stride;
    bee += layer * layer_stride;
    wpre += layer * layer_stride;
    int r0, r1, r2;
    bond_rows(code, r0, r1, r2);
    for (int h = threadIdx.x; h < H; h += blockDim.x)
        s_emb[h] = __ldg(bond_tables + r0 * H + h) + __ldg(bond_tables + r1 * H + h) + __ldg(bond_tables + r2 * H + h);
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float acc = 0.f;
        const float* wrow = wee + (int64_t)f * H;
        for (int h = 0; h < H; ++h) acc = fmaf(__ldg(wrow + h), s_emb[h], acc);
        acc += __ldg(bee + f);
        s_enc[f] = acc;
        enc[((int64_t)layer * GNN_NUM_BOND_CODES + code) * F + f] = acc;
    }
    __syncthreads();
    const int C = towers * F;
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
        int t = j / F, f = j - t * F;
        const float* wrow = wpre + t * wpre_tower_stride + (int64_t)f * 3 * F + 2 * F;
        float acc = 0.f;
        for (int k = 0; k < F; ++k) acc = fmaf(__ldg(wrow + k), s_enc[k], acc);
        tab[((int64_t)layer * GNN_NUM_BOND_CODES + code) * C + j] = acc;
    }
}

extern "C" int gnn_bond_table_fwd(const float* bond_tables, const float* wee, const float* bee, const float* wpre,
                                  int64_t wpre_tower_stride, int64_t layer_stride, int num_layers, int hidden,
                                  int towers, float* enc, float* tab, void* stream) {
    if (hidden < 1 || towers < 1 || num_layers < 1) return GNN_ERR_BAD_ARG;
    dim3 grid(GNN_NUM_BOND_CODES, num_layers);
    bond_table_fwd_kernel<<<grid, 128, 2 * hidden * sizeof(float), gnn_stream(stream)>>>(
        bond_tables, wee, bee, wpre, wpre_tower_stride, layer_stride, hidden, towers, enc, tab);
    return gnn_launch_status();
}

// backward, kernel A: grid (60, L): denc[l,c,:] and the embedding-table gradient
__global__ void __launch_bounds__(128)
bond_table_bwd_codes_kernel(const float* __restrict__ dtab, const float* __restrict__ wee,
                            const float* __restrict__ wpre, int64_t wpre_tower_stride, int64_t layer_stride, int H,
                            int towers, float* __restrict__ denc, float* dbond_tables) {
    extern __shared__ float sm[];
    const int F = H, C = towers * F;
    float* s_dtab = sm;        // [C]
    float* s_denc = sm + C;    // [F]
    const int code = blockIdx.x, layer = blockIdx.y;
    wee += layer * layer_stride;
    wpre += layer * layer_stride;
    for (int j = threadIdx.x; j < C; j += blockDim.x)
        s_dtab[j] = __ldg(dtab + ((int64_t)layer * GNN_NUM_BOND_CODES + code) * C + j);
    __syncthreads();
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        float acc = 0.f;
        for (int t = 0; t < towers; ++t) {
            const float* w = wpre + t * wpre_tower_stride + 2 * F + k;
            for (int f = 0; f < F; ++f) acc = fmaf(s_dtab[t * F + f], __ldg(w + (int64_t)f * 3 * F), acc);
        }
        s_denc[k] = acc;
        denc[((int64_t)layer * GNN_NUM_BOND_CODES + code) * F + k] = acc;
    }
    __syncthreads();
    int r0, r1, r2;
    bond_rows(code, r0, r1, r2);
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float acc = 0.f;
        for (int f = 0; f < F; ++f) acc = fmaf(s_denc[f], __ldg(wee + (int64_t)f * H + h), acc);
        atomicAdd(dbond_tables + r0 * H + h, acc);
        atomicAdd(dbond_tables + r1 * H + h, acc);
        atomicAdd(dbond_tables + r2 * H + h, acc);
    }
}
// kernel B: grid (T*F, L), block over k: dWpre_t[f, 2F+k] += sum_c dtab[c,tF+f] enc[c,k]
__global__ void bond_table_bwd_wpre_kernel(const float* __restrict__ dtab, const float* __restrict__ enc,
                                           int64_t wpre_tower_stride, int64_t layer_stride, int H, int towers,
                                           float* dwpre) {
    const int F = H, C = towers * F;
    const int j = blockIdx.x, layer = blockIdx.y;
    const int t = j / F, f = j - t * F;
    float* drow = dwpre + layer * layer_stride + t * wpre_tower_stride + (int64_t)f * 3 * F + 2 * F;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < GNN_NUM_BOND_CODES; ++c)
            acc = fmaf(__ldg(dtab + ((int64_t)layer * GNN_NUM_BOND_CODES + c) * C + j),
                       __ldg(enc + ((int64_t)layer * GNN_NUM_BOND_CODES + c) * F + k), acc);
        drow[k] += acc;
    }
}
// kernel C: grid (F, L), block over h: dWee[f,h] += sum_c denc[c,f] emb[c,h]; dbee[f] += sum_c denc[c,f]
__global__ void bond_table_bwd_wee_kernel(const float* __restrict__ denc, const float* __restrict__ bond_tables,
                                          int64_t layer_stride, int H, float* dwee, float* dbee) {
    const int F = H;
    const int f = blockIdx.x, layer = blockIdx.y;
    const float* de = denc + (int64_t)layer * GNN_NUM_BOND_CODES * F + f;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < GNN_NUM_BOND_CODES; ++c) {
            int r0, r1, r2;
            bond_rows(c, r0, r1, r2);
            float e = __ldg(bond_tables + r0 * H + h) + __ldg(bond_tables + r1 * H + h) + __ldg(bond_tables + r2 * H + h);
            acc = fmaf(__ldg(de + (int64_t)c * F), e, acc);
        }
        dwee[layer * layer_stride + (int64_t)f * H + h] += acc;
    }
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int c = 0; c < GNN_NUM_BOND_CODES; ++c) acc += __ldg(de + (int64_t)c * F);
        dbee[layer * layer_stride + f] += acc;
    }
}

extern "C" int gnn_bond_table_bwd(const float* dtab, const float* bond_tables, const float* wee, const float* wpre,
                                  int64_t wpre_tower_stride, int64_t layer_stride, int num_layers, const float* enc,
                                  int hidden, int towers, float* denc_ws, float* dbond_tables, float* dwee,
                                  float* dbee, float* dwpre, void* stream) {
    if (hidden < 1 || towers < 1 || num_layers < 1) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    const int F = hidden, C = towers * F;
    dim3 g1(GNN_NUM_BOND_CODES, num_layers);
    bond_table_bwd_codes_kernel<<<g1, 128, (C + F) * sizeof(float), st>>>(dtab, wee, wpre, wpre_tower_stride,
                                                                          layer_stride, hidden, towers, denc_ws,
                                                                          dbond_tables);
    dim3 g2(C, num_layers);
    bond_table_bwd_wpre_kernel<<<g2, 64, 0, st>>>(dtab, enc, wpre_tower_stride, layer_stride, hidden, towers, dwpre);
    dim3 g3(F, num_layers);
    bond_table_bwd_wee_kernel<<<g3, 64, 0, st>>>(denc_ws, bond_tables, layer_stride, hidden, dwee, dbee);
    return gnn_launch_status();
}
