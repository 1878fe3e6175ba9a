"""Hand-made worlds for branches the generated maps never reach (sparse_rail_generator makes neither
dead ends nor switch-free rail cycles): a one-way loop with a dead-end spur.  Transition masks use the
reference's encoding (core/grid/grid4.py:39-50): four nibbles for the heading N,E,S,W, each with exit bits
N,E,S,W."""
import numpy as np

V, HZ = 0x8020, 0x0401                       # straight north-south / east-west
TL, TR, BR, BL = 0x4002, 0x1200, 0x0810, 0x0048   # corners of a loop
DEAD_END_S = 0x0080                          # entered heading south, the only way is back north


def loop_world(n_agents=4, T=60):
    """7x8 map: a rectangular loop (rows 1..4, cols 1..6) and a spur (2,3)-(3,3) hanging from the top edge.
    Heading east through (1,3) only continues east (the switch is unusable from that side), so a branch walk
    that goes round clockwise comes back to its first state without meeting a switch: the `visited` branch of
    _explore_branch (treeobs.cpp:476-481).  Heading west through (1,3) is a real switch (west or south)."""
    H, W = 7, 8
    g = np.zeros((H, W), np.uint16)
    for c in range(2, 6):
        g[1, c] = HZ
        g[4, c] = HZ
    for r in range(2, 4):
        g[r, 1] = V
        g[r, 6] = V
    g[1, 1], g[1, 6], g[4, 6], g[4, 1] = TL, TR, BR, BL
    # (1,3): heading N (from the spur) -> E; heading E -> E; heading W -> W or S
    g[1, 3] = (0b0100 << 12) | (0b0100 << 8) | (0b0011 << 0)
    g[2, 3] = V
    g[3, 3] = DEAD_END_S
    init = [(3, 3), (1, 5), (4, 2), (2, 6), (3, 1), (4, 5)][:n_agents]
    idir = [0, 1, 3, 2, 0, 3][:n_agents]     # clockwise headings on the loop; the spur train heads north
    # targets: on the loop, and the dead end (unreachable for clockwise trains -> infinite distance, walks cycle)
    tgt = [(4, 4), (3, 3), (1, 2), (3, 3), (4, 3), (2, 1)][:n_agents]
    speed = [1.0, 0.5, 1.0, 0.25, 0.33, 1.0][:n_agents]
    return dict(H=H, W=W, N=n_agents, T=T, grid=g, init_pos=np.array(init, np.int16), init_dir=np.array(idir, np.uint8),
                target=np.array(tgt, np.int16), speed=np.array(speed, np.float64),
                earliest=np.array([0, 1, 2, 0, 3, 1][:n_agents], np.int32),
                latest=np.array([40, 45, 50, 55, 58, 59][:n_agents], np.int32))
