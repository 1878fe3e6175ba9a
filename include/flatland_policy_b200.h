/* flatland_policy_b200.h — C ABI of the batched policy forward pass: the consumer that
 * follows the step/observation path immediately, evaluated for all E*N agents of a lock-step batch from the
 * observation tensors where fl_observe left them (include/flatland_b200.h), without a host round trip.
 *
 * Replaces, in the reference (paths relative to the reference repository root):
 *   solution/nn/net_tree.py:73-98   Network.forward (attr MLP, 3 x MultiheadAttention block, actor / critic heads)
 *   solution/nn/TreeLSTM.py:34-154  TreeLSTM.forward / _run_lstm (child-sum style Tree-LSTM over the 31-node trees)
 *   solution/eval_env.py:76         forest[forest == inf] = -1
 *   solution/plfActor.py:15-44      Actor.get_actions / _choose_action (soft choice, generator re-seeded with 42)
 * The reference runs these through torch on the CPU with batch 1 (plfActor.py:18-20); there is no native
 * interface to bind, so the boundary is this C ABI under the Python class `BatchedActor`
 * (flatland-marl_b200/policy.py) that keeps `Actor.get_actions`' meaning.
 *
 * Arithmetic: bf16 operands, fp32 accumulation on the 5th-generation tensor cores (tcgen05.mma, accumulators in
 * TMEM); gates, GELU, softmax in fp32.  Tolerance against the fp32 reference is stated in tests/test_gpu_policy.py.
 *
 * Conventions as in flatland_b200.h: plain device pointers, `stream` is a cudaStream_t passed as void*, return
 * 0 or an FlStatus / CUDA error code, no allocation inside the library.
 */
#ifndef FLATLAND_POLICY_B200_H
#define FLATLAND_POLICY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FL_POLICY_ABI_VERSION 1
#define FL_POLICY_HIDDEN 128 /* NetworkConfig.hidden_sz = tree_embedding_sz (solution/impl_config.py:23-25) */
#define FL_POLICY_EMB 256
#define FL_POLICY_HEADS 4
#define FL_POLICY_LAYERS 3
#define FL_POLICY_MAX_LEVELS 11 /* node_order of a 31-node ternary tree is at most 10 */

/* Device-resident parameters.  Matrices are bf16 (raw uint16), row-major [out][in] exactly as nn.Linear stores
 * them, `in` zero-padded to the stated width; biases are fp32.  flatland-marl_b200/policy.py packs them from a
 * reference state_dict. */
typedef struct FlPolicyWeights {
    /* Tree-LSTM.  The gate biases ride in the matrix product: the node-feature operand carries a constant 1 in column 12
     * and column 12 of tree_wiou / column 140 of tree_ufwf holds the bias.  Rows of the sigmoid gates (i: 0..127, o:
     * 128..255 of tree_uiou / tree_wiou, and all of tree_ufwf) are stored multiplied by 1/2, because the kernels
     * evaluate sigmoid(2z) as 0.5 tanh(z) + 0.5 (one MUFU operation). */
    const uint16_t *tree_uiou;  /* [384][384]  tree_lstm.U_iou.weight (i, o rows x 1/2) */
    const uint16_t *tree_wiou;  /* [384][16]   tree_lstm.W_iou.weight | col 12: W_iou.bias (i, o rows x 1/2) */
    const uint16_t *tree_wc;    /* [128][384]  tree_lstm.W_c.weight */
    const uint16_t *tree_ufwf;  /* [128][144]  (tree_lstm.U_f.weight | tree_lstm.W_f.weight | col 140: W_f.bias) x 1/2 */
    const float *tree_b_iou;    /* [384] unscaled, kept for reference (not read by the kernels) */
    const float *tree_b_c;      /* [128] */
    const float *tree_b_f;      /* [128] unscaled, kept for reference (not read by the kernels) */
    const uint16_t *attr_w[4];  /* [256][128 (83 used)], [256][256], [256][256], [128][256]  attr_embedding.{0,2,4,6} */
    const float *attr_b[4];
    const uint16_t *tf_wqkv[FL_POLICY_LAYERS]; /* [768][256] transformer.l.attention.in_proj_weight */
    const float *tf_bqkv[FL_POLICY_LAYERS];
    const uint16_t *tf_wo[FL_POLICY_LAYERS];   /* [256][256] attention.out_proj.weight (kept for reference, not read: folded below) */
    const float *tf_bo[FL_POLICY_LAYERS];
    const uint16_t *tf_wm[FL_POLICY_LAYERS];   /* [256][512] att_mlp.0.weight with the out-projection folded in:
                                                  [W1 | W2 Wo], so that the layer reads cat(input, attention heads) */
    const float *tf_bm[FL_POLICY_LAYERS];      /* att_mlp.0.bias + W2 out_proj.bias */
    const uint16_t *head_w1;    /* [512][512] rows 0..255 actor_net.0.weight, rows 256..511 critic_net.0.weight */
    const float *head_b1;       /* [512] */
    const uint16_t *head_w2a;   /* [128][256] actor_net.2.weight */
    const uint16_t *head_w2c;   /* [128][256] critic_net.2.weight */
    const float *head_b2;       /* [256] actor_net.2.bias | critic_net.2.bias */
    const float *head_w3;       /* [6][128] fp32: rows 0..4 actor_net.4.weight, row 5 critic_net.4.weight */
    const float *head_b3;       /* [6] */
} FlPolicyWeights;

int fl_policy_abi_version(void);

/* Bytes of device scratch fl_policy_forward needs for n_agents_total = E*N agents (256-byte aligned blocks). */
size_t fl_policy_workspace_bytes(int64_t n_agents_total);

/* Network.forward for E environments of N agents each.  Inputs are the tensors fl_observe wrote
 * (d_agent_attr [E][N][83] f32, d_forest [E][N][31][12] f32 with +inf still in place, d_adjacency [E][N][30][3],
 * d_node_order [E][N][31], d_edge_order unused: edge order is node_order[parent]).
 * Outputs: d_logits [E][N][5] f32 (net_tree.py:94-96 worker_action), d_value [E] f32 (critic mean over agents,
 * net_tree.py:107-110).  d_workspace: fl_policy_workspace_bytes(E*N) bytes. */
int fl_policy_forward(const FlPolicyWeights *w, void *d_workspace, size_t workspace_bytes, int64_t E, int64_t N,
                      const float *d_agent_attr, const float *d_forest, const int32_t *d_adjacency,
                      const int32_t *d_node_order, float *d_logits, float *d_value, void *stream);

/* The same forward pass in the reference's own arithmetic: fp32 operands and accumulation on the CUDA cores, exact erf-GELU /
 * sigmoid / tanh, the reference's unfolded parameters (csrc/policy/policy_f32.cuh).  d_weights: FL_POLICY_F32_TENSORS device
 * pointers to the fp32 tensors of the reference state_dict in its registration order (solution/nn/net_tree.py:33-72,
 * TreeLSTM.py:12-32; flatland-marl_b200/policy_weights.py:weight_spec lists names and shapes).  Logits and values agree with
 * the reference network to summation-order rounding (tests: 1e-4 absolute); about ten times slower than the tensor-core
 * path.  Synchronises `stream` once (per-level node counts). */
#define FL_POLICY_F32_TENSORS 46
size_t fl_policy_workspace_bytes_f32(int64_t n_agents_total);
int fl_policy_forward_f32(const float *const *d_weights, void *d_workspace, size_t workspace_bytes, int64_t E, int64_t N,
                          const float *d_agent_attr, const float *d_forest, const int32_t *d_adjacency,
                          const int32_t *d_node_order, float *d_logits, float *d_value, void *stream);

/* Actor._choose_action in "soft" mode for every agent (plfActor.py:27-44): masked softmax over the valid actions,
 * then numpy's choice with the generator re-seeded to 42, i.e. the first action whose cumulative probability
 * exceeds 0.3745401188473625; no valid action -> 0.  d_valid_actions [n][5] u8, d_actions [n] u8. */
int fl_policy_choose_actions(const float *d_logits, const uint8_t *d_valid_actions, uint8_t *d_actions,
                             int64_t n_agents_total, void *stream);

/* One dense layer on the tensor-core path, exposed for tests: C[M][ldc] (bf16) = act(A[M][K] . W[N][K]^T + bias),
 * K and N multiples of 64 / 128, act 0 = none, 1 = GELU(erf). */
int fl_policy_linear(const uint16_t *d_a, int64_t lda, const uint16_t *d_w, const float *d_bias, uint16_t *d_c,
                     int64_t ldc, int64_t M, int64_t N, int64_t K, int act, void *stream);

/* Tuning only: fl_policy_linear that also writes SM-clock timestamps of CTA (0,0) into d_clocks[128]
 * ([0] MMA warp start, [1] weights resident, [2] end, [3] kernel start; [8+4t..] MMA thread per tile t < 8: accumulator
 * free, first stage full, last stage full; [48+4t..] epilogue warp 0: accumulator full, read, stored; [80+t] producer
 * starts tile t). */
int fl_policy_linear_debug(const uint16_t *d_a, int64_t lda, const uint16_t *d_w, const float *d_bias, uint16_t *d_c,
                           int64_t ldc, int64_t M, int64_t N, int64_t K, int act, long long *d_clocks, void *stream);

/* Tuning only: while set (non-NULL), k_tree_leaf of later forwards writes SM-clock stamps of its CTA 0 into d_clocks[128]. */
void fl_policy_debug_clocks(long long *d_clocks);

uint64_t fl_policy_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
