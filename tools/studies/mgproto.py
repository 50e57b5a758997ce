# Throwaway prototype: compare MG-PCG variants on the production operator (isotropic K, dx=dy).
import numpy as np, sys, time

def synth_c(N, seed=1):
    L = np.sqrt(2.0); dx = L/N
    x = (np.arange(N)+0.5)*dx
    X, Y = np.meshgrid(x, x, indexing='ij')
    r = np.sqrt(X*X+Y*Y); th = np.arctan2(Y, X)
    r0 = L*(0.45+0.08*np.sin(16*th)); w = 4*dx
    c = 0.5*(1-np.tanh((r-r0)/w))
    # second injector at opposite corner (parallel config)
    X2, Y2 = L-X, L-Y
    r2 = np.sqrt(X2*X2+Y2*Y2); th2 = np.arctan2(Y2, X2)
    c2 = 0.5*(1-np.tanh((r2-L*(0.45+0.08*np.sin(16*th2)))/w))
    rng = np.random.default_rng(seed)
    c = np.clip(np.maximum(c, c2)+1e-3*(rng.random((N, N))-0.5), 0, 1)
    return c, dx

def wells_q(N, dx):
    # crude wells: quarter discs radius 0.05 L
    L = N*dx; r = 0.05*L; val = 4/(np.pi*r*r)
    x = (np.arange(N)+0.5)*dx
    X, Y = np.meshgrid(x, x, indexing='ij')
    q = np.zeros((N, N))
    q[(X**2+Y**2) < r*r] = val
    q[((L-X)**2+(L-Y)**2) < r*r] = val
    q[((L-X)**2+Y**2) < r*r] = -val
    q[(X**2+(L-Y)**2) < r*r] = -val
    q -= q.mean()   # compat (pin handles anyway)
    return q

def links_from_c(c, dx, nu, M=41.0):
    N1, N2 = c.shape
    m = M**0.25
    cg = np.pad(c, 1, mode='edge')  # mirror ghost 1
    mu = 41.0/(1-cg+m*cg)**4
    # AW[i,j] i in 0..N1 (faces), j in 0..N2-1 ; stored (N1+1, N2)
    Kx = 2/(mu[:-1, 1:-1]+mu[1:, 1:-1])     # (N1+1, N2)
    Ky = 2/(mu[1:-1, :-1]+mu[1:-1, 1:])     # (N1, N2+1)
    Kv = 4/(mu[:-1, :-1]+mu[1:, :-1]+mu[:-1, 1:]+mu[1:, 1:])  # (N1+1,N2+1) vertex (i-1/2,j-1/2)
    AW = nu*Kx/dx**2; AS = nu*Ky/dx**2; AV = (1-nu)*Kv/(2*dx**2)
    return fold(AW, AS, AV)

def fold(AW, AS, AV):
    AW = AW.copy(); AS = AS.copy(); AV = AV.copy()
    AW[:, 0] += AV[:, 0]; AW[:, -1] += AV[:, -1]
    AS[0, :] += AV[0, :]; AS[-1, :] += AV[-1, :]
    AW[0, :] = 0; AW[-1, :] = 0; AS[:, 0] = 0; AS[:, -1] = 0
    AV[0, :] = 0; AV[-1, :] = 0; AV[:, 0] = 0; AV[:, -1] = 0
    return AW, AS, AV

class Level:
    def __init__(s, AW, AS, AV, pin):
        s.AW, s.AS, s.AV = AW, AS, AV
        s.N1, s.N2 = AS.shape[0], AW.shape[1]
        s.diag = AW[:-1]+AW[1:]+AS[:, :-1]+AS[:, 1:]+AV[:-1, :-1]+AV[1:, :-1]+AV[:-1, 1:]+AV[1:, 1:]
        s.pin = pin
        I, J = np.meshgrid(np.arange(s.N1), np.arange(s.N2), indexing='ij')
        s.red = ((I+J) % 2 == 0); s.black = ~s.red
        s.col4 = [(I % 2 == a) & (J % 2 == b) for a, b in [(0, 0), (1, 1), (0, 1), (1, 0)]]
    def offsum(s, p):
        P = np.pad(p, 1)
        AW, AS, AV = s.AW, s.AS, s.AV
        return (AW[:-1]*P[:-2, 1:-1]+AW[1:]*P[2:, 1:-1]+AS[:, :-1]*P[1:-1, :-2]+AS[:, 1:]*P[1:-1, 2:]
                + AV[:-1, :-1]*P[:-2, :-2]+AV[1:, :-1]*P[2:, :-2]+AV[:-1, 1:]*P[:-2, 2:]+AV[1:, 1:]*P[2:, 2:])
    def matvec(s, p):   # SPD form: (-A) p = diag*p - offsum
        r = s.diag*p-s.offsum(p)
        if s.pin is not None: r[s.pin] = 0
        return r
    def smooth(s, p, f, order, ncol):
        masks = [s.red, s.black] if ncol == 2 else s.col4
        if order < 0: masks = masks[::-1]
        for mk in masks:
            pn = (f+s.offsum(p))/s.diag
            p = np.where(mk, pn, p)
            if s.pin is not None: p[s.pin] = 0
        return p

def coarsen(l, mode):
    AW, AS, AV = l.AW, l.AS, l.AV
    if mode == 'redisc':
        AWc = (AW[::2, 0::2]+AW[::2, 1::2])/8
        ASc = (AS[0::2, ::2]+AS[1::2, ::2])/8
        AVc = AV[::2, ::2]/4
    elif mode == 'series':
        # harmonic series of half,full,half fine links across the coarse face, mean over strips
        def ser(a, b, c_):
            with np.errstate(divide='ignore', invalid='ignore'):
                return 2/(0.5/a+1/b+0.5/c_)
        A = AW
        fullW = np.zeros_like(A[::2])
        Ap = np.pad(A, ((1, 1), (0, 0)), mode='edge')
        a = Ap[0:-2:2][: fullW.shape[0]]; b = A[::2]; c_ = Ap[2::2][: fullW.shape[0]]
        k = np.where(b > 0, ser(np.where(a > 0, a, b), b, np.where(c_ > 0, c_, b)), 0)
        AWc = (k[:, 0::2]+k[:, 1::2])/8
        A = AS
        Ap = np.pad(A, ((0, 0), (1, 1)), mode='edge')
        a = Ap[:, 0:-2:2][:, : A[:, ::2].shape[1]]; b = A[:, ::2]; c_ = Ap[:, 2::2][:, : b.shape[1]]
        k = np.where(b > 0, ser(np.where(a > 0, a, b), b, np.where(c_ > 0, c_, b)), 0)
        ASc = (k[0::2]+k[1::2])/8
        AVc = AV[::2, ::2]/4
    pin = None if l.pin is None else (l.pin[0]//2, l.pin[1]//2)
    return Level(AWc, ASc, AVc, pin)

def restrict(r, kind):
    if kind == 'avg':
        return 0.25*(r[0::2, 0::2]+r[1::2, 0::2]+r[0::2, 1::2]+r[1::2, 1::2])
    # adjoint of bilinear /4
    def r1(a, ax):
        a = np.moveaxis(a, ax, 0)
        n = a.shape[0]//2
        lo = a[0::2]; hi = a[1::2]
        out = 0.75*lo+0.75*hi
        out[1:] += 0.25*hi[:-1]      # fine 2I-1 (odd, below)
        out[:-1] += 0.25*lo[1:]      # fine 2I+2 (even, above)
        out[0] += 0.25*lo[0]         # boundary: weight that would leave the domain stays
        out[-1] += 0.25*hi[-1]
        return np.moveaxis(out/2, 0, ax)
    return r1(r1(r, 0), 1)

def prolong(e, kind):
    if kind == 'const':
        return np.repeat(np.repeat(e, 2, 0), 2, 1)
    def p1(a, ax):
        a = np.moveaxis(a, ax, 0)
        ap = np.pad(a, ((1, 1),)+((0, 0),)*(a.ndim-1), mode='edge')
        out = np.empty((2*a.shape[0],)+a.shape[1:])
        out[0::2] = 0.75*a+0.25*ap[:-2]
        out[1::2] = 0.75*a+0.25*ap[2:]
        return np.moveaxis(out, 0, ax)
    return p1(p1(e, 0), 1)

def vcycle(levels, li, f, cfg):
    l = levels[li]
    if li == len(levels)-1:
        p = np.zeros_like(f)
        for _ in range(cfg['ncoarse']):
            p = l.smooth(p, f, +1, cfg['ncol']); p = l.smooth(p, f, -1, cfg['ncol'])
        return p
    p = np.zeros_like(f)
    for _ in range(cfg['nsm']): p = l.smooth(p, f, +1, cfg['ncol'])
    r = f-l.matvec(p)
    if l.pin is not None: r[l.pin] = 0
    fc = restrict(r, cfg['R'])
    lc = levels[li+1]
    if lc.pin is not None: fc[lc.pin] = 0
    ec = vcycle(levels, li+1, fc, cfg)
    e = prolong(ec, cfg['P'])
    if l.pin is not None: e[l.pin] = 0
    p = p+cfg.get('omega', 1.0)*e
    for _ in range(cfg['nsm']): p = l.smooth(p, f, -1, cfg['ncol'])
    return p

def pcg(levels, b, cfg, tol=1e-14, maxit=200):
    A = levels[0]
    x = np.zeros_like(b); r = b.copy()
    if A.pin is not None: r[A.pin] = 0
    bb = (b*b).sum()
    z = vcycle(levels, 0, r, cfg); d = z.copy(); gamma = (r*z).sum()
    hist = []
    for it in range(1, maxit+1):
        s = A.matvec(d); alpha = gamma/(s*d).sum()
        x += alpha*d; r -= alpha*s
        rr = (r*r).sum(); hist.append(np.sqrt(rr/bb))
        if rr/bb < tol*tol: break
        z = vcycle(levels, 0, r, cfg); g2 = (r*z).sum(); beta = g2/gamma; gamma = g2
        d = z+beta*d
    true = np.sqrt(((b-A.matvec(x))**2).sum()/bb)
    return x, it, hist, true

if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    c, dx = synth_c(N); q = wells_q(N, dx)
    AW, AS, AV = links_from_c(c, dx, 2/3)
    for cfg in CFGS:
        pass
    for cfg in [
        dict(ncol=2, R='avg', P='const', coarsen='redisc', nsm=1, ncoarse=8),
        dict(ncol=2, R='adj', P='bilin', coarsen='redisc', nsm=1, ncoarse=8),
        dict(ncol=4, R='adj', P='bilin', coarsen='redisc', nsm=1, ncoarse=8),
        dict(ncol=2, R='avg', P='bilin', coarsen='redisc', nsm=1, ncoarse=8),
        dict(ncol=2, R='adj', P='bilin', coarsen='series', nsm=1, ncoarse=8),
        dict(ncol=2, R='avg', P='const', coarsen='series', nsm=1, ncoarse=8),
    ]:
        levels = [Level(AW, AS, AV, (1, 1))]
        while levels[-1].N1 % 2 == 0 and levels[-1].N1 > 4:
            levels.append(coarsen(levels[-1], cfg['coarsen']))
        b = q.copy()
        t = time.time()
        x, it, hist, true = pcg(levels, b, cfg)
        i8 = next((k+1 for k, h in enumerate(hist) if h < 1e-8), None)
        print(N, cfg, 'its(1e-14)=', it, 'its(1e-8)=', i8, 'true=%.2e' % true, 'rho=%.3f' % ((hist[-1]/hist[0])**(1/(len(hist)-1))), '%.1fs' % (time.time()-t), flush=True)

# ---- second-generation helpers: sigma-pin, variable smoothing, cycle shape
class LevelS(Level):
    def __init__(s, AW, AS, AV, pin, sigma=None):
        super().__init__(AW, AS, AV, pin if sigma is None else None)
        s.spin = pin; s.sigma = sigma
        if sigma is not None:
            s.diag = s.diag.copy(); s.diag[pin] += sigma

def link_sum_siblings(l):
    i, j = l.spin
    bi, bj = (i//2)*2, (j//2)*2
    tot = 0.0
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di == 0 and dj == 0: continue
            ii, jj = i+di, j+dj
            if not (bi <= ii < bi+2 and bj <= jj < bj+2): continue
            if dj == 0: tot += l.AW[max(i, ii), j]
            elif di == 0: tot += l.AS[i, max(j, jj)]
            else: tot += l.AV[max(i, ii), max(j, jj)]
    return tot

def build(AW, AS, AV, cfg, fac=1/8, pin=(1, 1), nmin=4):
    levels = [LevelS(AW, AS, AV, pin)]
    sigma = None
    while levels[-1].N1 % 2 == 0 and levels[-1].N2 % 2 == 0 and min(levels[-1].N1, levels[-1].N2) > nmin:
        l = levels[-1]
        lc = coarsen(Level(l.AW, l.AS, l.AV, None), cfg['coarsen'])
        pinc = (l.spin[0]//2, l.spin[1]//2)
        sigma = fac*link_sum_siblings(l) if sigma is None else sigma/4
        levels.append(LevelS(lc.AW, lc.AS, lc.AV, pinc, sigma))
    return levels

def cycle(levels, li, f, cfg):
    l = levels[li]
    nsm = cfg['nsm'] if li == 0 else cfg.get('nsm_c', cfg['nsm'])
    if li == len(levels)-1:
        p = np.zeros_like(f)
        for _ in range(cfg['ncoarse']):
            p = l.smooth(p, f, +1, cfg['ncol']); p = l.smooth(p, f, -1, cfg['ncol'])
        return p
    p = np.zeros_like(f)
    for _ in range(nsm): p = l.smooth(p, f, +1, cfg['ncol'])
    gam = cfg.get('gamma', 1) if li >= cfg.get('gamma_from', 1) else 1
    for g in range(gam):
        r = f-l.matvec(p)
        fc = restrict(r, cfg['R'])
        ec = cycle(levels, li+1, fc, cfg)
        e = prolong(ec, cfg['P'])
        if l.pin is not None: e[l.pin] = 0
        p = p+e
        if g < gam-1:
            # symmetric W-cycle needs smoothing between coarse visits
            for _ in range(nsm): p = l.smooth(p, f, -1, cfg['ncol'])
            for _ in range(nsm): p = l.smooth(p, f, +1, cfg['ncol'])
    for _ in range(nsm): p = l.smooth(p, f, -1, cfg['ncol'])
    return p
