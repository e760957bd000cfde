/* gnnepcsaft_b200 — C ABI of the B200 (sm_100a) PNAPCSAFT message-passing path.
 *
 * The reference (wildsonbbl/GNN-ePC-SAFT) has no native boundary: every op of
 * gnnepcsaft/train/models.py:105-135 is a PyTorch / PyG / OGB library call.  Each entry
 * point below replaces one of those library-call sites (cited per function) and is what a
 * ctypes/cffi binding on the reference side binds.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host
 *  - the caller owns all memory; nothing is allocated inside; workspaces are sized by the
 *    *_workspace_bytes() queries
 *  - launches are enqueued on `stream` (a cudaStream_t passed as void*) and return
 *    immediately; no global state and no environment variables are read, so calls are
 *    thread-safe per stream (the only cache is the SM count per device ordinal)
 *  - return value: 0 = ok, >0 = cudaError_t of the launch, <0 = argument error
 *    (GNN_ERR_*); gnn_error_string() renders either
 *  - float tensors are fp32 row-major; `ld*` are leading dimensions in elements
 */
#ifndef GNNEPCSAFT_B200_H
#define GNNEPCSAFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNN_ERR_BAD_ARG (-1)
#define GNN_ERR_TOO_LARGE (-2)
#define GNN_ERR_UNSUPPORTED (-3)
#define GNN_ERR_WORKSPACE (-4)

/* bits of the device-side input error word (`err_flag` arguments): malformed inputs on which the reference's
 * torch indexing would raise IndexError are reported here instead of being silently clamped */
#define GNN_INPUT_ERR_EDGE_INDEX 1     /* edge_index outside [0, N) */
#define GNN_INPUT_ERR_BOND_FEATURE 2   /* edge_attr outside the OGB bond feature ranges */
#define GNN_INPUT_ERR_BATCH_RANGE 4    /* batch id outside [0, B) */
#define GNN_INPUT_ERR_BATCH_UNSORTED 8 /* batch vector not sorted (PyG Batch always is) */
#define GNN_INPUT_ERR_ATOM_FEATURE 16  /* x outside the OGB atom feature ranges */

#define GNN_NUM_ATOM_FEATS 9
#define GNN_NUM_BOND_FEATS 3
#define GNN_NUM_BOND_CODES 60 /* 5*6*2, code = a0*12 + a1*2 + a2; self-loop = code 0 */

const char* gnn_error_string(int code);
int gnn_abi_version(void);
/* device properties seen by the library: out[0]=sm count, out[1]=cc major, out[2]=cc minor */
int gnn_device_info(int32_t* out_host);

/* ---------------------------------------------------------------- graph structure
 * Replaces: PyG add_self_loops (models.py:118-121) and the per-layer scatter index
 * handling inside MessagePassing/scatter (models.py:128 -> PNAConv.propagate).
 * Builds, on device, the destination-grouped CSR (stable: original edge order inside a
 * segment, the implicit self-loop last, matching PyG's cat([edges, loops])) and the
 * source-grouped list of CSR slots used by the backward pass.
 *   rowptr   int32 [N+1]   col_src int32 [E']  col_code uint8 [E']  col_eid int32 [E']
 *   srcptr   int32 [N+1]   src_slot int32 [E']      E' = E + (self_loops ? N : 0)
 * col_eid[s] is the index into the PyG-order edge list (loops are E+i).
 * err_flag (int32[1], may be NULL) is set non-zero if an index or bond feature is out of
 * range (the reference would raise an IndexError): GNN_INPUT_ERR_* bits, cleared on entry. */
int64_t gnn_build_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int gnn_build_csr(const int64_t* edge_index, const int64_t* edge_attr, int64_t num_nodes,
                  int64_t num_edges, int self_loops, int32_t* rowptr, int32_t* col_src,
                  uint8_t* col_code, int32_t* col_eid, int32_t* srcptr, int32_t* src_slot,
                  int32_t* err_flag, void* workspace, int64_t workspace_bytes, void* stream);

/* batch vector (int64, sorted, PyG `Batch.batch`) -> int32 node->graph map and graph_ptr
 * [B+1].  Replaces the index handling of global_add_pool (models.py:133).  err_flag (may be
 * NULL) receives GNN_INPUT_ERR_BATCH_RANGE / _UNSORTED bits (OR-ed in, not cleared). */
int gnn_graph_ptr_from_batch(const int64_t* batch, int64_t num_nodes, int64_t num_graphs,
                             int32_t* node_graph, int32_t* graph_ptr, int32_t* err_flag,
                             void* stream);

/* single molecule (`data.batch is None`, reference models.py:116 / PyG global_add_pool): graph_ptr = [0, N] */
int gnn_graph_ptr_single(int64_t num_nodes, int32_t* graph_ptr, void* stream);

/* Device collate. Replaces PyG DataLoader/Collater/Batch.from_data_list (train.py:74-79)
 * over an InMemoryDataset-format packed store (graphdataset.py:48,137): concatenated
 * tensors with graph-local edge_index plus slices.  For the selected graph ids writes the
 * PyG-layout int64 batch (x, edge_index incremented, edge_attr, batch, ptr) bit-exactly.
 *   ids int64 [B]; node_slices/edge_slices int64 [G+1]
 *   out_ptr int64 [B+1], out_eptr int64 [B+1] (edge offsets, scratch/output) */
int64_t gnn_collate_workspace_bytes(int64_t num_ids);
int gnn_collate_offsets(const int64_t* ids, int64_t num_ids, const int64_t* node_slices,
                        const int64_t* edge_slices, int64_t* out_ptr, int64_t* out_eptr,
                        void* workspace, int64_t workspace_bytes, void* stream);
int gnn_collate_gather(const int64_t* ids, int64_t num_ids, const int64_t* node_slices,
                       const int64_t* edge_slices, const int64_t* store_x,
                       const int64_t* store_edge_index, int64_t store_num_edges,
                       const int64_t* store_edge_attr, const int64_t* out_ptr,
                       const int64_t* out_eptr, int64_t total_nodes, int64_t total_edges,
                       int64_t* x, int64_t* edge_index, int64_t* edge_attr, int64_t* batch,
                       void* stream);

/* Compact device store -> batch (dataset residency).  The store holds ONE destination CSR over the
 * whole dataset (gnn_build_csr on the concatenated graphs, once): node_off / slot_off int32 [G+1] are the node and
 * CSR-slot ranges of every graph, store_x uint8 [sumN, 9] (gnn_atom_features_to_u8), store_rowptr int32 [sumN+1],
 * store_col_src int32 (store-global node ids) / store_col_code uint8 [sumE'], store_para float [G, P] (may be NULL).
 * For graph ids int64 [B] it writes what the model kernels consume, with no sort and no scan over edges:
 *   x uint8 [Np, 9], rowptr int32 [Np+1], col_src int32 [Ep], col_code uint8 [Ep], node_graph int32 [Np],
 *   graph_ptr int32 [cap+2], para float [(cap+1) P], valid int32 [4] = {N, B, E', 0}
 * cap = cap_graphs >= B, Np = pad_nodes >= N and Ep = pad_slots >= E' (else GNN_INPUT_ERR_EDGE_INDEX in err_flag):
 * graphs B .. cap-1 are empty, the node / slot surplus becomes graph `cap` (isolated padding nodes holding the surplus
 * slots as self-edges), so every batch of a (cap, Np, Ep) shape bucket issues identical launches and one captured
 * CUDA graph serves the bucket.  Exact sizes give an empty padding graph.  own_surplus = 0: the padding nodes have no
 * in-edges and the surplus CSR slots are unused capacity (rowptr[Np] = E'); enough for models whose kernels are all
 * node-indexed (pre_layers == 1, the reference's configs) - the captured graph then does not depend on Ep at all.
 * own_surplus = 1: every slot of the padded arrays belongs to a node (per-slot tensors of the pre_layers > 1 path).
 * Replaces PyG DataLoader / Collater (train.py:74-79) + add_self_loops (models.py:118-121) for store-resident data. */
int64_t gnn_compact_collate_workspace_bytes(int64_t num_ids);
int gnn_compact_collate(const int64_t* ids, int64_t num_ids, int64_t cap_graphs, int64_t num_store_graphs,
                        const int32_t* node_off, const int32_t* slot_off, const uint8_t* store_x,
                        const int32_t* store_rowptr, const int32_t* store_col_src,
                        const uint8_t* store_col_code, const float* store_para, int num_para,
                        int64_t pad_nodes, int64_t pad_slots, int own_surplus, uint8_t* x, int32_t* rowptr,
                        int32_t* col_src, uint8_t* col_code, int32_t* node_graph, int32_t* graph_ptr,
                        float* para, int32_t* valid, int32_t* err_flag, void* workspace,
                        int64_t workspace_bytes, void* stream);
/* int64 [N, 9] OGB atom features -> uint8 (values outside [0, 255]: GNN_INPUT_ERR_ATOM_FEATURE) */
int gnn_atom_features_to_u8(const int64_t* x, int64_t num_nodes, uint8_t* out, int32_t* err_flag,
                            void* stream);

/* ---------------------------------------------------------------- encoders
 * OGB AtomEncoder (models.py:65,122): out[i,:] = sum_k table[row_off[k] + x[i,k], :].
 * `tables` is the 9 embedding weights stacked along rows (174 rows), row_off int32[9]. */
int gnn_atom_embed_fwd(const int64_t* x, int64_t num_nodes, const float* tables,
                       const int32_t* row_off, int32_t total_rows, int hidden, float* out,
                       int32_t* err_flag, void* stream);
/* the same on compact uint8 features (device store / compact batches; every OGB atom feature is < 119) */
int gnn_atom_embed_u8_fwd(const uint8_t* x, int64_t num_nodes, const float* tables,
                          const int32_t* row_off, int32_t total_rows, int hidden, float* out,
                          int32_t* err_flag, void* stream);
int gnn_atom_embed_u8_bwd(const uint8_t* x, int64_t num_nodes, const float* dout,
                          const int32_t* row_off, int32_t total_rows, int hidden, float* dtables,
                          void* stream);
/* d tables[row] += sum_i dout[i,:] (tables gradient must be zeroed by the caller) */
int gnn_atom_embed_bwd(const int64_t* x, int64_t num_nodes, const float* dout,
                       const int32_t* row_off, int32_t total_rows, int hidden, float* dtables,
                       void* stream);

/* OGB BondEncoder + PNAConv.edge_encoder + the bond third of pre_nns[t][0]
 * (models.py:66,123; PyG pna_conv.py message()).  With pre_layers == 1 the per-edge pre-NN
 * is linear, so the bond contribution depends only on the 60 possible bond codes:
 *   emb[c,:]  = sum_k bond_tables[boff[k] + c_k, :]                       [60,H]
 *   enc[c,:]  = emb[c,:] . Wee^T + bee                                    [60,F]
 *   tab[c, t*F+f] = sum_k Wpre_t[f, 2F+k] * enc[c,k]                      [60,T*F]
 * wpre points at layer 0's pre_nns[0][0].weight, towers are `wpre_tower_stride` floats apart,
 * ld = 3F; all `num_layers` conv layers are done in one launch, the parameters of layer l sit
 * `l*layer_stride` floats after layer 0's (flat parameter arena), outputs are [L,60,F] /
 * [L,60,T*F].  enc is also returned (needed by backward). */
int gnn_bond_table_fwd(const float* bond_tables, const float* wee, const float* bee,
                       const float* wpre, int64_t wpre_tower_stride, int64_t layer_stride,
                       int num_layers, int hidden, int towers, float* enc, float* tab,
                       void* stream);
/* backward of the above given dtab [L,60,T*F]: accumulates into dbond_tables [13,H],
 * dwee [F,H], dbee [F], dwpre (bond third only; same addressing as the parameters).
 * denc_ws: scratch [L,60,F]. */
int gnn_bond_table_bwd(const float* dtab, const float* bond_tables, const float* wee,
                       const float* wpre, int64_t wpre_tower_stride, int64_t layer_stride,
                       int num_layers, const float* enc, int hidden, int towers, float* denc_ws,
                       float* dbond_tables, float* dwee, float* dbee, float* dwpre, void* stream);

/* ---------------------------------------------------------------- fused message + aggregate
 * Replaces MessagePassing._collect/message + MultiAggregation(mean,min,max,std)
 * (PyG pna_conv.py, nn/aggr/{multi,basic,scaler}.py; called from models.py:128).
 * Inputs: P [N, 2C] = [Pdst | Psrc] node projections (C = towers*F), tab [60,C].
 *   q_e = Psrc[src_e] + tab[code_e];   m_e = Pdst[dst_e] + q_e
 * Output A [N, towers, 4, F]: mean, min, max (of m) and std (of q; shift invariant), with
 * empty segments -> 0 and std = sqrt(max(var,1e-5)) masked to 0 when <= sqrt(1e-5).
 * One warp per destination node, single pass, variance from sums shifted by the first message. */
int gnn_pna_aggregate_fwd(const float* P, const float* tab, const int32_t* rowptr,
                          const int32_t* col_src, const uint8_t* col_code, int64_t num_nodes,
                          int towers, int F, float* A, void* stream);
/* Backward: gA [N,towers,4,F] -> dP[:, 0:C] (dst part), dq [E',C] per CSR slot and
 * dtab_accum [60,C] (+=, caller zeroes).  min/max gradients are split evenly over ties
 * (torch scatter_reduce amin/amax semantics).  `zeroed_workspace` holds partial bond-code
 * tables; it must be all-zero on entry and is all-zero again on exit.
 * dq == NULL: the per-slot gradients are not materialised; dP[:, C:2C] (source part) is
 * cleared and receives dq_e at row src_e through 128-bit atomic adds (summation order
 * not fixed), which replaces the gnn_scatter_src_bwd pass. */
int64_t gnn_pna_aggregate_bwd_workspace_bytes(int towers, int F);
int gnn_pna_aggregate_bwd(const float* gA, const float* P, const float* tab, const float* A,
                          const int32_t* rowptr, const int32_t* col_src,
                          const uint8_t* col_code, int64_t num_nodes, int towers, int F,
                          float* dP, float* dq, float* dtab_accum, void* zeroed_workspace,
                          int64_t workspace_bytes, void* stream);
/* dP[:, C:2C][j] = sum over out-edges of j of dq[slot] (deterministic order) */
int gnn_scatter_src_bwd(const float* dq, const int32_t* srcptr, const int32_t* src_slot,
                        int64_t num_nodes, int C, float* dP, void* stream);

/* Folded destination projection (pre_layers == 1, every node has an in-edge, i.e. self-loops on):
 *   mean/min/max(m) = Wdst x + mean/min/max(q), std(m) = std(q), and post_nns[t][0] is linear, so the Wdst x
 * term moves into the x columns of post_nns:  Vx[t][s] = [s==0] Wpost_t[:, :F] + (Wmean+Wmin+Wmax)_{t,s} . Wdst_t
 * (s = identity / amplification / attenuation group).  The aggregate kernels then read source projections
 * only (Psrc [N, C], bias of pre_nns included) and emit the aggregates of q.
 *   gnn_pna_fold_fwd : Vx [L,T,3,F_out,F] from the parameters of all L conv layers in one launch (wpost/wpre
 *       point at tower 0 of layer 0; layer l sits l*layer_stride floats further, as in gnn_bond_table_fwd)
 *   gnn_pna_fold_bwd : dVx -> dwpost (x columns and the mean/min/max columns of every group), dwpre (dst third); +=
 *   gnn_pna_aggregate_src_fwd/bwd : as gnn_pna_aggregate_fwd/bwd with P = Psrc [N, C] and no destination term;
 *       bwd writes dq [E',C] per CSR slot, or with dq == NULL accumulates atomically into dPsrc [N, C]
 *   gnn_slot_sum_by_source : out[j, 0:C] (row stride ldo) = sum over out-edges of j of dq[slot] */
int gnn_pna_fold_fwd(const float* wpost, int64_t wpost_tower_stride, const float* wpre,
                     int64_t wpre_tower_stride, int64_t layer_stride, int num_layers, int F, int F_out,
                     int towers, float* Vx, void* stream);
int gnn_pna_fold_bwd(const float* dVx, const float* wpost, int64_t wpost_tower_stride, const float* wpre,
                     int64_t wpre_tower_stride, int64_t layer_stride, int num_layers, int F, int F_out,
                     int towers, float* dwpost, float* dwpre, void* stream);
int gnn_pna_aggregate_src_fwd(const float* Psrc, const float* tab, const int32_t* rowptr,
                              const int32_t* col_src, const uint8_t* col_code, int64_t num_nodes,
                              int towers, int F, float* A, void* stream);
int gnn_pna_aggregate_src_bwd(const float* gA, const float* Psrc, const float* tab, const float* A,
                              const int32_t* rowptr, const int32_t* col_src, const uint8_t* col_code,
                              int64_t num_nodes, int towers, int F, float* dPsrc, float* dq,
                              float* dtab_accum, void* zeroed_workspace, int64_t workspace_bytes,
                              void* stream);
int gnn_slot_sum_by_source(const float* dq, const int32_t* srcptr, const int32_t* src_slot,
                           int64_t num_nodes, int C, float* out, int64_t ldo, void* stream);

/* General message path (pre_layers > 1): per-edge hidden h_e = relu(Pdst[dst]+Psrc[src]+tab[code])
 * in CSR slot order, [E', C]. */
int gnn_edge_hidden_fwd(const float* P, const float* tab, const int32_t* rowptr,
                        const int32_t* col_src, const uint8_t* col_code, int64_t num_nodes,
                        int C, int relu, float* hidden_out, void* stream);
/* Aggregate materialised per-slot messages m [E', C] (CSR order) -> A. */
int gnn_pna_aggregate_msg_fwd(const float* m, const int32_t* rowptr, int64_t num_nodes,
                              int towers, int F, float* A, void* stream);
int gnn_pna_aggregate_msg_bwd(const float* gA, const float* m, const float* A,
                              const int32_t* rowptr, int64_t num_nodes, int towers, int F,
                              float* dm, void* stream);
/* backward of gnn_edge_hidden_fwd: dh [E',C] (already masked by relu) -> dP dst half (+=0,
 * written), dq copy is dh itself; accumulates dtab. */
int gnn_edge_hidden_bwd(const float* dh, const int32_t* rowptr, const uint8_t* col_code,
                        int64_t num_nodes, int C, float* dP, float* dtab_accum, void* stream);

/* ---------------------------------------------------------------- dense stages (fp32 SIMT)
 * Replace torch.nn.Linear / PyG Linear (addmm) at pna_conv.py pre_nns/post_nns/lin and the
 * MLP head (models.py:84-103).  All are  C[M,N] (+)= op(A) . op(B) with row-major storage.
 *
 * gnn_linear_fwd:  Y[M,N] = act(X[M,K] . W[N,K]^T + bias)        (act: 0 none, 1 relu)
 * gnn_linear_dgrad: dX[M,K] (+)= dY[M,N] . W[N,K]                (accumulate: 0/1)
 * gnn_linear_wgrad: dW[N,K] += dY[M,N]^T . X[M,K],  db[N] += colsum(dY)  (db may be NULL)
 * gnn_relu_bwd:     dY[i] = Y_act[i] > 0 ? dY[i] : 0   (ReLU between Linear layers of pre/post nns)
 */
int gnn_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                   int64_t M, int N, int K, int act, float* Y, int64_t ldy, void* stream);
int gnn_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t M,
                     int N, int K, int accumulate, float* dX, int64_t lddx, void* stream);
int gnn_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t M,
                     int N, int K, float* dW, int64_t lddw, float* db, void* stream);
int gnn_relu_bwd(float* dY, const float* Y_act, int64_t numel, void* stream);

/* Degree scalers of DegreeScalerAggregation (PyG nn/aggr/scaler.py): scal[i] = (amp_i, att_i),
 * amp = log(d+1)/avg_deg_log, att = avg_deg_log/log(max(d,1)+1), d = in-degree incl. loops. */
int gnn_pna_scalers(const int32_t* rowptr, const float* avg_deg_log, int64_t num_nodes,
                    float* scal, void* stream);

/* PNAConv post_nns[t][0] over the virtual concatenation [x | A_t | amp*A_t | att*A_t]
 * (K = 13F) - the [N,T,12F]/[N,T,13F] tensors of DegreeScalerAggregation / PNAConv.forward
 * are never materialised.
 *   fwd:   O[:, t*F_out:(t+1)*F_out] = act(cat . Wpost_t^T + bpost_t),  Wpost_t [F_out,13F]
 *   dgrad: dx[N,F] += sum_t dO_t . Wpost_t[:, :F];  gA[:, t, :, :] = folded scaler gradient
 *   wgrad: dWpost_t += dO_t^T . cat ;  dbpost_t += colsum(dO_t)
 * wpost/bpost point at tower 0; towers are *_tower_stride floats apart. */
int gnn_pna_post_fwd(const float* x, const float* A, const float* scal, const float* wpost,
                     int64_t wpost_tower_stride, const float* bpost, int64_t bpost_tower_stride,
                     int64_t num_nodes, int F, int F_out, int towers, int act, float* O,
                     void* stream);
int gnn_pna_post_dgrad(const float* dO, const float* scal, const float* wpost,
                       int64_t wpost_tower_stride, int64_t num_nodes, int F, int F_out, int towers,
                       float* dx, float* gA, void* stream);
int gnn_pna_post_wgrad(const float* dO, const float* x, const float* A, const float* scal,
                       int64_t num_nodes, int F, int F_out, int towers, float* dwpost,
                       int64_t wpost_tower_stride, float* dbpost, int64_t bpost_tower_stride,
                       void* stream);

/* ---------------------------------------------------------------- dense stages (tcgen05 / TMEM)
 * The same contractions on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
 * `precision`: 0 = TF32x3 (fp32 operands split into tf32 hi+lo, 3 MMAs: fp32-level accuracy, the
 * parity path), 2 = BF16 operands, fp32 accumulate (the reference's own
 * torch.set_float32_matmul_precision("medium"), train.py:319).  Semantics equal the gnn_linear_* /
 * gnn_pna_post_* entry points above (bias gradients are column sums taken by the operand producers of the
 * weight-gradient kernels).  Operands must be 128-bit
 * accessible (16-byte aligned, leading dimensions and K / N multiples of 4), else GNN_ERR_UNSUPPORTED.
 * gnn_tc_pna_post_dgrad: dx may be NULL (then only gA is produced and the x part goes through
 * gnn_tc_pna_dx_dgrad).
 * pack_ws (optional, 128-byte aligned, may be NULL): scratch in which the weight operand is first packed
 * as the exact shared-memory tile image (swizzled, hi/lo planes) so the GEMM fetches it with TMA bulk
 * copies instead of re-converting it for every 128-row tile; 4 MiB covers every shape of the default model, the
 * wide one (hidden 256) needs 8 MiB for its post GEMM (the host passes 32 MiB).  Too small -> unpacked (slower) path.
 * pack_ws_bytes < 0 (inference with constant weights): pack_ws (|pack_ws_bytes| bytes) already holds the image that an
 * earlier call of the same entry point with the same weights and shapes packed there - the pack kernel is skipped;
 * GNN_ERR_WORKSPACE if the image cannot have fitted.
 * Row count 0 with pack_ws_bytes > 0: only the weight image is packed (no activation pointer is read; NULL is fine),
 * for the same call with the real row count and -pack_ws_bytes later.  The host issues these packs of a captured
 * training step on a second stream, where they depend on the weights only and leave the step's critical path. */
int gnn_tc_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                      int64_t M, int N, int K, int act, float* Y, int64_t ldy, int precision,
                      void* pack_ws, int64_t pack_ws_bytes, void* stream);
int gnn_tc_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t M, int N,
                        int K, int accumulate, float* dX, int64_t lddx, int precision, void* pack_ws,
                        int64_t pack_ws_bytes, void* stream);
int gnn_tc_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t M, int N,
                        int K, float* dW, int64_t lddw, float* db, int precision, void* stream);
int gnn_tc_pna_post_fwd(const float* x, const float* A, const float* scal, const float* wpost,
                        int64_t wpost_tower_stride, const float* bpost, int64_t bpost_tower_stride,
                        int64_t num_nodes, int F, int F_out, int towers, int act, float* O,
                        int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream);
int gnn_tc_pna_post_dgrad(const float* dO, const float* scal, const float* wpost,
                          int64_t wpost_tower_stride, int64_t num_nodes, int F, int F_out, int towers,
                          float* dx, float* gA, int precision, void* pack_ws, int64_t pack_ws_bytes,
                          void* stream);
int gnn_tc_pna_post_wgrad(const float* dO, const float* x, const float* A, const float* scal,
                          int64_t num_nodes, int F, int F_out, int towers, float* dwpost,
                          int64_t wpost_tower_stride, float* dbpost, int64_t bpost_tower_stride,
                          int precision, void* stream);

/* Fused projection stage of PNAConv (pre_nns[t][0], first two thirds of its input), all towers in one
 * launch.  wpre / bpre point at tower 0 of the layer, towers are *_tower_stride floats apart.
 *   proj_fwd : P[N, 2C] = [x Wdst_t^T + b_t | x Wsrc_t^T]                    (C = towers*F)
 *   dx_dgrad : dx (+)= dP . [Wdst | Wsrc] + sum_t dO_t . Wpost_t[:, :F]      (all node-feature gradients of a layer)
 *   proj_wgrad: dWpre_t[:, :2F] += dP^T x ;  dbpre_t += colsum(dPdst_t) */
int gnn_tc_pna_proj_fwd(const float* x, const float* wpre, int64_t wpre_tower_stride, const float* bpre,
                        int64_t bpre_tower_stride, int64_t num_nodes, int F, int towers, float* P,
                        int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream);
int gnn_tc_pna_dx_dgrad(const float* dP, const float* dO, const float* wpre, int64_t wpre_tower_stride,
                        const float* wpost, int64_t wpost_tower_stride, int64_t num_nodes, int F, int F_out,
                        int towers, int accumulate, float* dx, int precision, void* pack_ws,
                        int64_t pack_ws_bytes, void* stream);
int gnn_tc_pna_proj_wgrad(const float* dP, const float* x, int64_t num_nodes, int F, int towers,
                          float* dwpre, int64_t wpre_tower_stride, float* dbpre, int64_t bpre_tower_stride,
                          int precision, void* stream);

/* Tensor-core stages of the folded destination projection (see gnn_pna_fold_fwd):
 *   projsrc_fwd  : Psrc[N, C] = x Wsrc_t^T + b_t
 *   postf_fwd    : O_t = act(sum_s scal_s ([x | A^q_t] . [Vx_{t,s} | Wagg_{t,s}]^T) + bpost_t)
 *   postf_wgrad  : aggregate columns -> dwpost, dbpost as gnn_tc_pna_post_wgrad; x columns of the three scaler
 *                  groups -> dVx [T,3,F_out,F] (+=, caller zeroes)
 *   dxf_dgrad    : dx (+)= dPsrc . Wsrc + sum_{s,t} scal_s dO_t . Vx_{t,s}
 *   projsrc_wgrad: dWpre_t[:, F:2F] += dPsrc_t^T x ;  dbpre_t += colsum(dPsrc_t) */
int gnn_tc_pna_projsrc_fwd(const float* x, const float* wpre, int64_t wpre_tower_stride, const float* bpre,
                           int64_t bpre_tower_stride, int64_t num_nodes, int F, int towers, float* Psrc,
                           int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream);
int gnn_tc_pna_postf_fwd(const float* x, const float* A, const float* scal, const float* Vx,
                         const float* wpost, int64_t wpost_tower_stride, const float* bpost,
                         int64_t bpost_tower_stride, int64_t num_nodes, int F, int F_out, int towers, int act,
                         float* O, int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream);
/* postf_fwd with the aggregates computed inside the GEMM (replaces the pair gnn_pna_aggregate_src_fwd +
 * gnn_tc_pna_postf_fwd, i.e. PyG MessagePassing._collect / message / MultiAggregation + post_nns of PNAConv at
 * reference models.py:128, in ONE kernel): the operand producers gather Psrc[src] + tab[code] over each node's CSR
 * segment, reduce mean / min / max / std in registers (the arithmetic of gnn_pna_aggregate_src_fwd, bit-identical)
 * and write them straight into the shared-memory operand tiles, so the [N, T, 4, F] aggregates are not read from
 * HBM and - with A_out == NULL (inference) - never written either.  A_out != NULL (training) receives them for the
 * backward pass.  precision must be 0 (TF32x3); F % 8 == 0. */
int gnn_tc_pna_postf_agg_fwd(const float* x, const float* Psrc, const float* tab, const int32_t* rowptr,
                             const int32_t* col_src, const uint8_t* col_code, const float* scal,
                             const float* Vx, const float* wpost, int64_t wpost_tower_stride,
                             const float* bpost, int64_t bpost_tower_stride, int64_t num_nodes, int F,
                             int F_out, int towers, int act, float* O, float* A_out, int precision,
                             void* pack_ws, int64_t pack_ws_bytes, void* stream);
int gnn_tc_pna_postf_wgrad(const float* dO, const float* x, const float* A, const float* scal,
                           int64_t num_nodes, int F, int F_out, int towers, float* dwpost,
                           int64_t wpost_tower_stride, float* dbpost, int64_t bpost_tower_stride, float* dVx,
                           int precision, void* stream);
int gnn_tc_pna_dxf_dgrad(const float* dPsrc, const float* dO, const float* scal, const float* wpre,
                         int64_t wpre_tower_stride, const float* Vx, int64_t num_nodes, int F, int F_out,
                         int towers, int accumulate, float* dx, int precision, void* pack_ws,
                         int64_t pack_ws_bytes, void* stream);
int gnn_tc_pna_projsrc_wgrad(const float* dPsrc, const float* x, int64_t num_nodes, int F, int towers,
                             float* dwpre, int64_t wpre_tower_stride, float* dbpre, int64_t bpre_tower_stride,
                             int precision, void* stream);

/* ---------------------------------------------------------------- BatchNorm / ReLU / skip
 * Replaces PyG BatchNorm / torch BatchNorm1d + F.relu + skip (models.py:128-131, head BN).
 * Training: batch statistics (biased var for normalisation, unbiased for running_var,
 * momentum 0.1), num_batches_tracked += 1.  stats = [mean[C] | rstd[C]] is written for
 * backward.  Eval: stats derived from running_mean / running_var.
 * valid_rows (device int32[1], may be NULL): padded batches - only the first *valid_rows rows are real (the rest
 * belongs to the padding graph of a shape bucket, see gnn_compact_collate): statistics, backward sums and their
 * 1/n use that count and the gradient rows of the padding are written as zero.  The count is read on the device
 * so that one captured CUDA graph serves every batch of a bucket. */
int64_t gnn_bn_workspace_bytes(int64_t rows, int C);
int gnn_bn_stats(const float* Y, int64_t rows, int C, int training, float eps, float momentum,
                 float* running_mean, float* running_var, int64_t* num_batches_tracked,
                 float* stats, const int32_t* valid_rows, void* workspace, int64_t workspace_bytes,
                 void* stream);
/* out = relu(gamma*(Y-mean)*rstd+beta) [+ skip]  (relu: 0/1, skip may be NULL) */
int gnn_bn_act_fwd(const float* Y, const float* stats, const float* gamma, const float* beta,
                   const float* skip, int64_t rows, int C, int relu, float* out, void* stream);
/* dY from dOut; also dgamma[C] += , dbeta[C] += .  training=0 uses the eval formula. */
int gnn_bn_act_bwd(const float* dOut, const float* Y, const float* stats, const float* gamma,
                   const float* beta, int64_t rows, int C, int relu, int training, float* dY,
                   float* dgamma, float* dbeta, const int32_t* valid_rows, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* Inference fast path: eval-mode BatchNorm folded into the preceding Linear, then Linear + ReLU + skip add in
 * one tensor-core GEMM (replaces lin -> BatchNorm -> relu -> x + x_prev of models.py:128-131 and the head's
 * Linear -> BatchNorm1d -> ReLU blocks when the model is in eval()).
 *   gnn_bn_fold_linear   : Wf[n,:] = W[n,:] gamma[n] rstd[n];  bf[n] = (bias[n] - mean[n]) gamma[n] rstd[n] + beta[n]
 *                          (Wf dense [N,K], rstd = 1/sqrt(running_var + eps); bias may be NULL)
 *   gnn_tc_linear_skip_fwd: Y = act(X W^T + bias) + skip   (skip may be NULL; N, K multiples of 4) */
int gnn_bn_fold_linear(const float* W, int64_t ldw, const float* bias, const float* gamma, const float* beta,
                       const float* running_mean, const float* running_var, float eps, int N, int K,
                       float* Wf, float* bf, void* stream);
int gnn_tc_linear_skip_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                           int64_t M, int N, int K, int act, const float* skip, int64_t ldskip, float* Y,
                           int64_t ldy, int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream);

/* ---------------------------------------------------------------- inference: the whole PNAConv stack per tile
 * Store-resident inference of the default architecture (F = 64, F_out = 32, 2 towers, folded destination projection,
 * post_layers == 1; other shapes: GNN_ERR_UNSUPPORTED).  Replaces, for ALL layers at once, the chain
 * gnn_tc_pna_projsrc_fwd -> gnn_tc_pna_postf_agg_fwd -> gnn_bn_fold_linear + gnn_tc_linear_skip_fwd, i.e.
 * PNAConv + BatchNorm (running statistics) + ReLU + skip of reference models.py:126-131: one persistent CTA keeps a
 * tile of <= 128 nodes made of whole graphs (no edge leaves a graph) in shared memory - node features, source
 * projections, CSR slice - and runs every layer on it with tcgen05 MMAs (TF32x3); only packed weights are re-read.
 *   gnn_tc_infer_pack   : packed weight images of all layers (wpre / wpost point at tower 0 of layer 0, layer l sits
 *                         l*layer_stride floats further; Vx [L,T,3,F_out,F] from gnn_pna_fold_fwd; wlin_folded
 *                         [L,H,H] from gnn_bn_fold_linear)
 *   gnn_tc_infer_tiles  : greedy packing of consecutive graphs (graph_ptr [G+1], only the first *valid_graphs when
 *                         given) into tiles = int2 (first node, end node) with <= 128 nodes and <= 1024 CSR slots;
 *                         tiles needs room for one tile per graph; a graph that fits no tile sets err_flag
 *   gnn_tc_pna_stack_infer : x_out [N, H] = the conv stack applied to x0 [N, H]; tab [L,60,T*F] bond tables
 *                         (gnn_bond_table_fwd), bpre / bpost = bias of pre_nns[0][0] / post_nns[0][0] of layer 0
 *                         (+ tower / layer strides), blin [L,H] folded bias, avg_deg_log [L]; rows of nodes outside
 *                         every tile are left untouched */
int64_t gnn_tc_infer_image_bytes(int num_layers);
int gnn_tc_infer_pack(const float* wpre, int64_t wpre_tower_stride, const float* Vx, const float* wpost,
                      int64_t wpost_tower_stride, int64_t layer_stride, const float* wlin_folded,
                      int num_layers, int F, int F_out, int towers, void* images, int64_t image_bytes,
                      void* stream);
int gnn_tc_infer_tiles(const int32_t* graph_ptr, const int32_t* rowptr, int64_t num_graphs,
                       const int32_t* valid_graphs, int32_t* tiles, int64_t max_tiles, int32_t* num_tiles,
                       int32_t* err_flag, void* stream);
int gnn_tc_pna_stack_infer(const float* x0, const int32_t* rowptr, const int32_t* col_src,
                           const uint8_t* col_code, const int32_t* tiles, const int32_t* num_tiles,
                           const void* images, const float* tab, const float* bpre, int64_t bpre_tower_stride,
                           int64_t bpre_layer_stride, const float* bpost, int64_t bpost_tower_stride,
                           int64_t bpost_layer_stride, const float* blin, const float* avg_deg_log,
                           int num_layers, int F, int F_out, int towers, int skip, float* x_out,
                           void* stream);

/* ---------------------------------------------------------------- pooling, loss, optimiser
 * global_add_pool (models.py:133): out[g,:] = sum_{i in graph g} x[i,:]; graph_ptr int32 [B+1] */
int gnn_segment_pool_fwd(const float* x, const int32_t* graph_ptr, int64_t num_graphs, int C,
                         float* out, void* stream);
int gnn_segment_pool_bwd(const float* dout, const int32_t* node_graph, int64_t num_nodes, int C,
                         int accumulate, float* dx, void* stream);
/* torchmetrics MAPE (models.py:194): loss[0] = mean(|p-t| / max(|t|,1.17e-6));
 * dpred = d loss / d pred * grad_scale.  valid_rows (device int32[1], may be NULL) with row_len = num_para:
 * padded batches - only the first *valid_rows rows of pred / target are real graphs; the mean runs over those and
 * the gradient rows of the padding graph are written as zero (see gnn_compact_collate). */
int gnn_mape_fwd_bwd(const float* pred, const float* target, int64_t numel, float grad_scale,
                     float* loss, float* dpred, const int32_t* valid_rows, int row_len, void* stream);
/* AdamW(amsgrad=True) + CosineAnnealingWarmRestarts(T_0) stepped per step
 * (models.py:162-188).  step_count int64[1] on device is read (t = step_count before this
 * step), the caller bumps it with gnn_step_increment.  grad_scale multiplies grads (1/world). */
int gnn_adamw_amsgrad_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           float* max_exp_avg_sq, int64_t numel, const int64_t* step_count,
                           float lr0, float beta1, float beta2, float eps, float weight_decay,
                           int64_t t0_restart, float grad_scale, void* stream);
int gnn_step_increment(int64_t* step_count, void* stream);
int gnn_fill_zero(void* ptr, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
