import sys, time, numpy as np
sys.path.insert(0, '/root/repo/tools/studies')
import mgproto as M
from vstudy import pcg  # noqa
def build_nopin(AW, AS, AV, cfg, nmin=4):
    levels = [M.Level(AW, AS, AV, None)]
    while levels[-1].N1 % 2 == 0 and min(levels[-1].N1, levels[-1].N2) > nmin:
        levels.append(M.coarsen(levels[-1], cfg['coarsen']))
    return levels
base = dict(ncol=2, R='adj', P='bilin', coarsen='redisc', nsm=1, ncoarse=8)
for N in (128, 256, 512, 1024):
    c, dx = M.synth_c(N); q = M.wells_q(N, dx)
    for name, cc in (('const', np.zeros_like(c)), ('synth', c)):
        AW, AS, AV = M.links_from_c(cc, dx, 2/3)
        it_pin = pcg(M.build(AW, AS, AV, base), q.copy(), base)
        it_nopin = pcg(build_nopin(AW, AS, AV, base), q.copy(), base)
        print(N, name, 'V its pinned', it_pin, 'singular/no pin', it_nopin, flush=True)
