"""Parameters of the reference policy network (solution/nn/net_tree.py:33-72 `Network`,
solution/nn/TreeLSTM.py:12-32 `TreeLSTM`) as a flat dict of numpy float32 arrays keyed by the
reference's own `state_dict` names, so that a checkpoint written by the reference
(`torch.save(net.state_dict())`, loaded in solution/plfActor.py:10-13) can be used as is.

The reference ships no trained checkpoint (solution/demo.py:43-52 names `policy/phase-III-*.pt`,
which are not in the repository), so tests and benchmarks use `init_weights(seed)`: uniform
nn.Linear-style initialisation bounds drawn from a numpy RandomState, which is
reproducible without torch's generator and without the reference on the machine.
"""
import numpy as np

HIDDEN = 128       # NetworkConfig.hidden_sz (solution/impl_config.py:23-25)
TREE = 128         # NetworkConfig.tree_embedding_sz
NODE_F = 12        # FeatureParserConfig.node_sz
ATTR_F = 83        # FeatureParserConfig.agent_attr
ACTIONS = 5
EMB = HIDDEN + TREE
HEADS = 4
N_TRANSFORMER = 3


def weight_spec():
    """[(state_dict name, shape, fan_in)] in the reference's registration order."""
    spec = [
        ("tree_lstm.W_iou.weight", (3 * TREE, NODE_F), NODE_F), ("tree_lstm.W_iou.bias", (3 * TREE,), NODE_F),
        ("tree_lstm.U_iou.weight", (3 * TREE, 3 * TREE), 3 * TREE),
        ("tree_lstm.W_c.weight", (TREE, 3 * TREE), 3 * TREE), ("tree_lstm.W_c.bias", (TREE,), 3 * TREE),
        ("tree_lstm.W_f.weight", (TREE, NODE_F), NODE_F), ("tree_lstm.W_f.bias", (TREE,), NODE_F),
        ("tree_lstm.U_f.weight", (TREE, TREE), TREE),
    ]
    dims = [(ATTR_F, 2 * HIDDEN), (2 * HIDDEN, 2 * HIDDEN), (2 * HIDDEN, 2 * HIDDEN), (2 * HIDDEN, HIDDEN)]
    for i, (k, n) in enumerate(dims):
        spec += [("attr_embedding.%d.weight" % (2 * i), (n, k), k), ("attr_embedding.%d.bias" % (2 * i), (n,), k)]
    for l in range(N_TRANSFORMER):
        p = "transformer.%d." % l
        spec += [(p + "attention.in_proj_weight", (3 * EMB, EMB), EMB), (p + "attention.in_proj_bias", (3 * EMB,), EMB),
                 (p + "attention.out_proj.weight", (EMB, EMB), EMB), (p + "attention.out_proj.bias", (EMB,), EMB),
                 (p + "att_mlp.0.weight", (EMB, 2 * EMB), 2 * EMB), (p + "att_mlp.0.bias", (EMB,), 2 * EMB)]
    for head, out in (("actor_net", ACTIONS), ("critic_net", 1)):
        dims = [(2 * EMB, 2 * HIDDEN), (2 * HIDDEN, HIDDEN), (HIDDEN, out)]
        for i, (k, n) in enumerate(dims):
            spec += [("%s.%d.weight" % (head, 2 * i), (n, k), k), ("%s.%d.bias" % (head, 2 * i), (n,), k)]
    return spec


def init_weights(seed=0, gain=2.5):
    """Uniform +-gain/sqrt(fan_in) matrices and +-1/sqrt(fan_in) biases.  gain 1 is nn.Linear's default; the
    default here (2.5, about Kaiming-uniform) keeps the signal alive through the 20-odd layers so that the
    outputs depend visibly on the observation — with gain 1 the logits are the biases to three digits, which
    would make a parity test blind."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, shape, fan_in in weight_spec():
        b = (gain if len(shape) > 1 else 1.0) / np.sqrt(fan_in)
        out[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
    return out


def check_weights(w):
    for name, shape, _ in weight_spec():
        if name not in w:
            raise KeyError("policy weights: missing %s" % name)
        if tuple(w[name].shape) != tuple(shape):
            raise ValueError("policy weights: %s has shape %s, expected %s" % (name, tuple(w[name].shape), shape))
    extra = sorted(set(w) - {name for name, _, _ in weight_spec()})
    if extra:
        raise KeyError("policy weights: unexpected keys %s (not a checkpoint of the reference Network)" % ", ".join(extra[:5]))
    return w


def load_weights(path):
    """A reference checkpoint (`torch.save(state_dict)`, plfActor.py:10-13) or an .npz with the same keys."""
    if str(path).endswith(".npz"):
        with np.load(path) as z:
            w = {k: np.asarray(z[k], np.float32) for k in z.files}
    else:
        import torch
        sd = torch.load(path, map_location="cpu", weights_only=True)
        w = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in sd.items()}
    return check_weights(w)
