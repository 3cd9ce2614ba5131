import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import test_gpu_training as T
from nsc_b200.training import CQTrainer
from nsc_b200 import codec
res_all, lsf_all = T.inputs(4, seed=95)
for sel in ([0], [1], [2], [3], [0, 1, 2, 3]):
    ocs, cm, cfg = T.make_models(2, -20.0)
    res_x, lsf = res_all[sel], lsf_all[sel]
    quan_w, ent_w, tau = [0., 0., 1.], [0., 0., 0.], 0.0
    tr = CQTrainer(cm, (60., 10., 10., tau), quan_w=quan_w, ent_w=ent_w)
    out = tr.loss_and_grads(torch.from_numpy(res_x).cuda(), torch.from_numpy(lsf).cuda(), tau=tau)
    flats, lsf_g, info, total = T.oracle_grads(ocs, cfg, -20.0, res_x, lsf, 1.0, (60., 10., 10.), quan_w, ent_w, tau, 2.0)
    g = tr.grads[1].cpu().numpy().astype(np.float64)
    tab = codec.layer_table(cfg)
    msg = []
    for li in (0, 1, 2, 3, 4):
        L = tab[li]; n = L.k * L.cin * L.cout
        a, b = g[L.offset:L.offset + n], flats[1][L.offset:L.offset + n]
        msg.append(f'L{li}: err {T.rel_l2(a, b):.1e} |ref| {np.linalg.norm(b):.3e}')
    print('frames', sel, 'time', info['time'], ' | '.join(msg))
