import sys, time, numpy as np
sys.path.insert(0, '/root/repo/tools/studies')
import mgproto as M
def pcg(levels, b, cfg, tol=1e-14, maxit=200):
    A = levels[0]
    x = np.zeros_like(b); r = b.copy()
    if A.pin is not None: r[A.pin] = 0
    bb = (b*b).sum()
    z = M.cycle(levels, 0, r, cfg); d = z.copy(); gamma = (r*z).sum()
    hist = []
    for it in range(1, maxit+1):
        s = A.matvec(d); alpha = gamma/(s*d).sum()
        x += alpha*d; r -= alpha*s
        rr = (r*r).sum(); hist.append(np.sqrt(rr/bb))
        if rr/bb < tol*tol: break
        z = M.cycle(levels, 0, r, cfg); g2 = (r*z).sum(); beta = g2/gamma; gamma = g2
        d = z+beta*d
    return it
if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    c, dx = M.synth_c(N); q = M.wells_q(N, dx)
    base = dict(ncol=2, R='adj', P='bilin', coarsen='redisc', nsm=1, ncoarse=8)
    for name, cc, nu, extra in [
        ('const nu=1 V', np.zeros_like(c), 1.0, {}),
        ('const nu=2/3 V', np.zeros_like(c), 2/3, {}),
        ('synth nu=1 V', c, 1.0, {}),
        ('synth nu=2/3 V', c, 2/3, {}),
        ('synth nu=2/3 V series', c, 2/3, dict(coarsen='series')),
        ('synth nu=2/3 V nsm=2', c, 2/3, dict(nsm=2)),
        ('synth nu=2/3 V nsm_c=2', c, 2/3, dict(nsm_c=2)),
        ('synth nu=2/3 V nsm_c=3', c, 2/3, dict(nsm_c=3)),
        ('synth nu=2/3 W from1', c, 2/3, dict(gamma=2, gamma_from=1)),
        ('synth nu=2/3 W from2', c, 2/3, dict(gamma=2, gamma_from=2)),
    ]:
        cfg = dict(base); cfg.update(extra)
        AW, AS, AV = M.links_from_c(cc, dx, nu)
        levels = M.build(AW, AS, AV, cfg)
        t = time.time()
        it = pcg(levels, q.copy(), cfg)
        print(N, name, 'its', it, '%.1fs' % (time.time()-t), flush=True)
