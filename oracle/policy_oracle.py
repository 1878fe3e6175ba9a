"""CPU restatement of the reference policy network — TEST INFRASTRUCTURE ONLY.

numpy float32 restatement of solution/nn/net_tree.py `Network.forward` (:73-98), solution/nn/TreeLSTM.py
`TreeLSTM.forward/_run_lstm` (:34-154), the observation clean-up of solution/eval_env.py:76
(`forest[forest == inf] = -1`) and the action choice of solution/plfActor.py:15-44.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; it is pinned against outputs of
the unmodified reference network recorded in tests/golden/policy_golden.npz (made by
tests/golden/make_policy_golden.py, checked by tests/test_policy_oracle.py).
"""
import numpy as np
from scipy.special import erf

F32 = np.float32


def gelu(x):  # nn.GELU() default: exact erf form
    x = x.astype(F32)
    return (x * F32(0.5) * (F32(1.0) + erf(x / F32(np.sqrt(2.0))).astype(F32))).astype(F32)


def sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x.astype(F32)))).astype(F32)


def linear(x, w, b=None):
    y = x.astype(F32) @ w.T.astype(F32)
    return (y + b).astype(F32) if b is not None else y.astype(F32)


def clean_forest(forest):
    """eval_env.py:76 — only +inf is replaced."""
    f = np.array(forest, dtype=F32, copy=True)
    f[f == np.inf] = -1
    return f


def tree_lstm(w, forest, adjacency, node_order, edge_order):
    """TreeLSTM.forward (TreeLSTM.py:34-57) after Network.modify_adjacency (net_tree.py:100-116).
    forest [B,N,31,12], adjacency [B,N,30,3], node_order [B,N,31], edge_order [B,N,30]; returns h of
    every node [B*N*31, 128]."""
    B, N, M, _ = forest.shape
    adj = adjacency.astype(np.int64).copy()
    tree_id = np.arange(B * N, dtype=np.int64).reshape(B, N, 1)
    adj[adj == -2] = -B * N * M
    adj[..., 0] += tree_id * M
    adj[..., 1] += tree_id * M
    adj[adj < 0] = -2
    x_all = forest.reshape(-1, forest.shape[-1]).astype(F32)
    adj = adj.reshape(-1, 3)
    no = node_order.reshape(-1)
    eo = edge_order.reshape(-1)
    T = w["tree_lstm.U_f.weight"].shape[0]
    h = np.zeros((x_all.shape[0], T), F32)
    c = np.zeros((x_all.shape[0], T), F32)
    for it in range(int(no.max()) + 1):
        nm = no == it
        em = eo == it
        x = x_all[nm]
        if it == 0:
            iou = linear(x, w["tree_lstm.W_iou.weight"], w["tree_lstm.W_iou.bias"])
        else:
            a = adj[em]
            par, chi = a[:, 0], a[:, 1]
            ch, cc = h[chi], c[chi]
            merge = ch.reshape(ch.shape[0] // 3, 3 * T)
            iou = linear(x, w["tree_lstm.W_iou.weight"], w["tree_lstm.W_iou.bias"]) + linear(merge, w["tree_lstm.U_iou.weight"])
        i, o, u = iou[:, :T], iou[:, T:2 * T], iou[:, 2 * T:]
        i, o, u = sigmoid(i), sigmoid(o), np.tanh(u).astype(F32)
        if it == 0:
            cn = i * u
        else:
            f = sigmoid(linear(x_all[par], w["tree_lstm.W_f.weight"], w["tree_lstm.W_f.bias"]) + linear(ch, w["tree_lstm.U_f.weight"]))
            fc = (f * cc).reshape(ch.shape[0] // 3, 3 * T)
            cn = i * u + linear(fc, w["tree_lstm.W_c.weight"], w["tree_lstm.W_c.bias"])
        c[nm] = cn
        h[nm] = o * np.tanh(cn).astype(F32)
    return h


def attention_block(w, p, x, heads=4):
    """Transformer.forward (net_tree.py:20-32): nn.MultiheadAttention over the agent axis, no mask, then
    GELU(Linear(cat(input, attention)))."""
    B, N, E = x.shape
    d = E // heads
    qkv = linear(x.reshape(B * N, E), w[p + "attention.in_proj_weight"], w[p + "attention.in_proj_bias"]).reshape(B, N, 3, heads, d)
    q = qkv[:, :, 0].transpose(0, 2, 1, 3) * F32(1.0 / np.sqrt(d))
    k = qkv[:, :, 1].transpose(0, 2, 1, 3)
    v = qkv[:, :, 2].transpose(0, 2, 1, 3)
    s = (q @ k.transpose(0, 1, 3, 2)).astype(F32)
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s).astype(F32)
    a = (e / e.sum(axis=-1, keepdims=True)).astype(F32)
    o = (a @ v).astype(F32).transpose(0, 2, 1, 3).reshape(B * N, E)
    o = linear(o, w[p + "attention.out_proj.weight"], w[p + "attention.out_proj.bias"])
    y = linear(np.concatenate([x.reshape(B * N, E), o], axis=1), w[p + "att_mlp.0.weight"], w[p + "att_mlp.0.bias"])
    return gelu(y).reshape(B, N, E)


def forward(w, agent_attr, forest, adjacency, node_order, edge_order):
    """Network.forward (net_tree.py:73-98): returns (logits [B,N,5], value [B]).  `forest` must already be
    cleaned (clean_forest)."""
    B, N, M, _ = forest.shape
    h = tree_lstm(w, forest, adjacency, node_order, edge_order)
    tree_emb = h.reshape(B, N, M, -1)[:, :, 0, :]
    a = agent_attr.reshape(B * N, -1).astype(F32)
    for i in range(4):
        a = gelu(linear(a, w["attr_embedding.%d.weight" % (2 * i)], w["attr_embedding.%d.bias" % (2 * i)]))
    emb = np.concatenate([a.reshape(B, N, -1), tree_emb], axis=2)
    att = emb
    for l in range(3):
        att = attention_block(w, "transformer.%d." % l, att)
    z = np.concatenate([emb, att], axis=-1).reshape(B * N, -1)

    def head(name):
        y = gelu(linear(z, w[name + ".0.weight"], w[name + ".0.bias"]))
        y = gelu(linear(y, w[name + ".2.weight"], w[name + ".2.bias"]))
        return linear(y, w[name + ".4.weight"], w[name + ".4.bias"])

    logits = head("actor_net").reshape(B, N, -1)
    value = head("critic_net").reshape(B, N).mean(axis=1)
    return logits, value.astype(F32)


CHOICE_U = 0.3745401188473625  # np.random.seed(42); np.random.random_sample() — the one draw np.random.choice makes


def choose_actions(logits, valid_actions):
    """Actor.get_actions / _choose_action, soft mode (plfActor.py:15-44): the generator is re-seeded with 42
    before every choice, so the choice is a deterministic function of the masked softmax."""
    logits = np.asarray(logits, F32)
    valid = np.asarray(valid_actions)
    out = np.zeros(logits.shape[:-1], np.uint8)
    for idx in np.ndindex(*logits.shape[:-1]):
        va = valid[idx]
        nz = va.nonzero()[0]
        if nz.size == 0:
            out[idx] = 0            # valid_actions = ones((1,5)) -> nonzero()[0] is all zeros -> action 0
            continue
        x = logits[idx][va != 0]
        e = np.exp(x - np.max(x))
        p = e / e.sum()
        np.random.seed(42)
        out[idx] = np.random.choice(nz, p=p)
    return out


def choice_margin(logits, valid_actions):
    """Distance of the sampling threshold from the nearest cdf step — tests skip agents whose margin is below
    the float tolerance of the compared implementation."""
    logits = np.asarray(logits, np.float64)
    valid = np.asarray(valid_actions)
    out = np.ones(logits.shape[:-1], np.float64)
    for idx in np.ndindex(*logits.shape[:-1]):
        va = valid[idx]
        if va.nonzero()[0].size == 0:
            continue
        x = logits[idx][va != 0]
        e = np.exp(x - x.max())
        cdf = np.cumsum(e / e.sum())
        out[idx] = np.min(np.abs(cdf[:-1] - CHOICE_U)) if cdf.size > 1 else 1.0
    return out
