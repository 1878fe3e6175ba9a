"""GPU check of the policy kernels: dense layers against torch, the whole forward against the numpy oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flatland_marl_b200 as fb  # noqa: E402
from flatland_marl_b200.policy import BatchedActor  # noqa: E402
from oracle import policy_oracle as po  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    actor = BatchedActor(None, seed=0)
    dev = actor.device
    torch.manual_seed(0)
    for (M, K, N, act) in [(128, 64, 128, 0), (128, 128, 128, 0), (300, 256, 256, 1), (51200, 256, 768, 0), (1000, 512, 256, 1), (77, 128, 384, 0)]:
        a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        out = actor.linear(a, w, b, act)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + b
        if act:
            ref = torch.nn.functional.gelu(ref)
        err = (out.float() - ref).abs().max().item()
        print("linear M=%d K=%d N=%d act=%d: max abs err %.4g (ref max %.3g)" % (M, K, N, act, err, ref.abs().max().item()), flush=True)
    with np.load(os.path.join(GOLD, "policy_golden.npz")) as z:
        gold = {k: z[k] for k in z.files}
    w = actor.weights
    cache = {}
    worst, agree, total = 0.0, 0, 0
    for k, (fixture, step) in enumerate(zip(gold["case_fixture"], gold["case_step"])):
        fixture = str(fixture)
        if fixture not in cache:
            with np.load(os.path.join(GOLD, fixture + ".npz")) as z:
                cache[fixture] = {kk: z[kk] for kk in z.files}
        g = cache[fixture]
        p = "obs%d_" % int(step)
        obs = {"agent_attr": g[p + "attr"], "forest": g[p + "forest"], "adjacency": g[p + "adjacency"], "node_order": g[p + "node_order"],
               "valid_actions": g[p + "valid_actions"]}
        dobs = {kk: torch.from_numpy(np.ascontiguousarray(v[None])).to(dev) for kk, v in obs.items()}
        logits, value = actor.forward(dobs)
        acts = actor.choose_actions(logits, dobs["valid_actions"])
        torch.cuda.synchronize()
        lg = logits[0].cpu().numpy()
        err = np.abs(lg - gold["logits_%d" % k]).max()
        verr = abs(float(value[0].item()) - float(gold["value_%d" % k][0]))
        worst = max(worst, err, verr)
        a = acts[0].cpu().numpy()
        agree += int((a == gold["actions_%d" % k]).sum())
        total += a.size
        # the device action choice applied to the device logits must equal the oracle's choice on the same logits
        same = (po.choose_actions(lg, obs["valid_actions"]) == a)
        margin = po.choice_margin(lg, obs["valid_actions"]) > 1e-5
        print("%s step %d N=%d: logits err %.4g value err %.4g actions==reference %.3f choice-kernel ok %s" %
              (fixture, int(step), lg.shape[0], err, verr, (a == gold["actions_%d" % k]).mean(), bool(same[margin].all())), flush=True)
    print("worst abs err %.4g; actions equal to the fp32 reference: %d / %d" % (worst, agree, total))
    print("policy launches:", actor.launch_count())


if __name__ == "__main__":
    main()
