"""Design study (dev tool): how many rays are active per march step, for plain sphere tracing and candidate accelerations."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import prior as P, scenes, sdf_oracle as O, trace_oracle as T

torch.set_num_threads(16)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prior = P.load_prior("assets/deepsdf_synth.pt")
K = scenes.intrinsics(size)
lat = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0)
pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0]))
o, d, rn = T.rays(K, size, size, pose)
inv = 1.0 / d
ta, tb = (T.BOX_LO - o) * inv, (T.BOX_HI - o) * inv
t0 = torch.minimum(ta, tb).max(dim=1)[0].clamp(min=0.0)
t1 = torch.maximum(ta, tb).min(dim=1)[0]

def f_of(idx, tau):
    x = o + tau[idx, None] * d[idx]
    with torch.no_grad():
        return O.decoder_forward(prior, torch.cat([lat.expand(idx.numel(), -1), x], 1)).squeeze(1)

def march(omega=1.0, near=5e-3, max_steps=64):
    active = t0 <= t1
    tau = t0.clone()
    prev_f = torch.zeros_like(tau); prev_step = torch.zeros_like(tau)
    nearset = torch.zeros_like(active)
    counts = []
    for step in range(max_steps):
        idx = active.nonzero().squeeze(1)
        counts.append(idx.numel())
        if idx.numel() == 0: break
        f = f_of(idx, tau)
        # over-relaxation fallback: if the unbounding spheres do not overlap, step back
        if omega > 1.0:
            bad = (prev_step[idx] > 0) & (f.abs() + prev_f[idx].abs() < prev_step[idx])
            # redo: go back to prev position + prev_f (plain step)
            tb_ = tau[idx] - prev_step[idx] + prev_f[idx]
            tau[idx[bad]] = tb_[bad]
            prev_step[idx[bad]] = 0
            good = ~bad
            idx = idx[good]; f = f[good]
        nr = f.abs() < near
        nearset[idx[nr]] = True
        active[idx[nr]] = False
        go = ~nr
        st = f[go] * omega
        prev_f[idx[go]] = f[go]; prev_step[idx[go]] = st if omega > 1.0 else 0
        tau[idx[go]] += st
        out = (tau[idx] > t1[idx]) | (tau[idx] < 0)
        active[idx[out & go]] = False
    return counts, int(nearset.sum()), tau, nearset

for om in (1.0, 1.3, 1.6):
    t = time.time()
    c, n, tau, ns = march(om)
    print(f"omega {om}: near {n}, evals {sum(c)}, steps-with-rays {sum(1 for x in c if x)}, {time.time()-t:.1f}s")
    print("  active per step:", c)

# ---- lattice-cached start -------------------------------------------------------------------------
Dn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
c0 = T.BOX_LO; hh = (T.BOX_HI - T.BOX_LO) / (Dn - 1)
ax = c0 + torch.arange(Dn, dtype=torch.float32) * hh
pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
with torch.no_grad():
    g = O.decoder_forward(prior, torch.cat([lat.expand(pts.shape[0], -1), pts], 1)).squeeze(1)
G = g.view(Dn, Dn, Dn)

def interp(x):
    u = ((x - c0) / hh).clamp(0, Dn - 1 - 1e-4)
    i = u.floor().long(); w = u - i
    out = 0
    for a in (0, 1):
        for b in (0, 1):
            for c in (0, 1):
                ww = (w[:, 0] if a else 1 - w[:, 0]) * (w[:, 1] if b else 1 - w[:, 1]) * (w[:, 2] if c else 1 - w[:, 2])
                out = out + ww * G[i[:, 0] + a, i[:, 1] + b, i[:, 2] + c]
    return out

# interpolation error on random points
xr = torch.rand(200000, 3) * (T.BOX_HI - c0) + c0
with torch.no_grad():
    fr = O.decoder_forward(prior, torch.cat([lat.expand(xr.shape[0], -1), xr], 1)).squeeze(1)
e = (interp(xr) - fr)
print("trilinear error: max", float(e.abs().max()), "p99.9", float(e.abs().quantile(0.999)), "near surface max", float(e[fr.abs() < 0.1].abs().max()))

def grid_start(delta=0.03, stop=0.06, max_steps=64):
    active = t0 <= t1
    tau = t0.clone()
    near = torch.zeros_like(active)
    for s in range(max_steps):
        idx = active.nonzero().squeeze(1)
        if idx.numel() == 0: break
        x = o + tau[idx, None] * d[idx]
        f = interp(x)
        nr = f < stop
        near[idx[nr]] = True; active[idx[nr]] = False
        go = ~nr
        tau[idx[go]] += (f[go] - delta)
        out = tau[idx] > t1[idx]
        active[idx[out & go]] = False
    return tau, near

def march2(tau, active, near=5e-3, max_steps=64, mode="plain"):
    tau = tau.clone(); active = active.clone()
    nearset = torch.zeros_like(active)
    prev_f = torch.zeros_like(tau); prev_step = torch.zeros_like(tau); om = torch.full_like(tau, 1.0)
    counts = []
    for step in range(max_steps):
        idx = active.nonzero().squeeze(1)
        counts.append(idx.numel())
        if idx.numel() == 0: break
        f = f_of(idx, tau)
        if mode == "auto":
            ps = prev_step[idx]
            bad = (ps > 0) & (f.abs() + prev_f[idx].abs() < ps)
            tau[idx[bad]] = (tau[idx] - ps + prev_f[idx])[bad]
            prev_step[idx[bad]] = 0; om[idx[bad]] = 1.0
            good = ~bad
            # slope estimate from the last two samples: m = (f - prev_f) / ps  (<= 0 when approaching)
            m = torch.where(ps > 0, (f - prev_f[idx]) / ps.clamp(min=1e-9), torch.full_like(f, -1.0))
            idx_g = idx[good]; f = f[good]; m = m[good]
            # planar assumption: distance along the ray to the surface f / (-m); over-relax towards it, capped
            w_ = (1.0 / (-m).clamp(min=0.25)).clamp(max=4.0) * 0.9
            w_ = torch.where(prev_step[idx_g] > 0, w_.clamp(min=1.0), torch.full_like(w_, 1.0))
            idx = idx_g
        else:
            w_ = torch.ones_like(f)
        nr = f.abs() < near
        nearset[idx[nr]] = True; active[idx[nr]] = False
        go = ~nr
        st = f[go] * w_[go]
        prev_f[idx[go]] = f[go]; prev_step[idx[go]] = st
        tau[idx[go]] += st
        out = (tau[idx] > t1[idx]) | (tau[idx] < 0)
        active[idx[out & go]] = False
    return counts, nearset, tau

c_ref, n_ref, tau_ref, ns_ref = march(1.0)
for delta, stop in ((0.03, 0.06), (0.02, 0.05), (0.04, 0.08)):
    tau_g, near_g = grid_start(delta, stop)
    for mode in ("plain", "auto"):
        c, ns, tau2 = march2(tau_g, near_g, mode=mode)
        both = ns & ns_ref
        print(f"grid start delta {delta} stop {stop} {mode}: rays to MLP {int(near_g.sum())}, evals {sum(c)}, steps {sum(1 for x in c if x)}, near {int(ns.sum())} (ref {n_ref}, common {int(both.sum())}), max |tau diff| on common {float((tau2-tau_ref)[both].abs().max()):.4f}")
        print("   ", c)

# ---- speculative K-sample march -------------------------------------------------------------------
def spec_march(tau_in, active_in, f0_in, m0_in, near=5e-3, c=1.6, rows_budget=16384, kmax=32, max_launch=40, verbose=True):
    tau = tau_in.clone(); active = active_in.clone()
    fh = f0_in.clone(); mh = m0_in.clone()      # predicted f at tau, slope estimate
    nearset = torch.zeros_like(active)
    log = []
    for it in range(max_launch):
        idx = active.nonzero().squeeze(1)
        n = idx.numel()
        if n == 0: break
        K = 1
        while K * 2 <= kmax and K * 2 * n <= rows_budget: K *= 2
        r = (1 + c * mh[idx]).clamp(0.3, 1.0)
        dt0 = c * fh[idx].abs().clamp(min=near)
        ks = torch.arange(K, dtype=torch.float32)
        geo = torch.where((1 - r[:, None]).abs() < 1e-4, ks[None, :].expand(n, K), (1 - r[:, None] ** ks[None, :]) / (1 - r[:, None]).clamp(min=1e-4))
        s = tau[idx, None] + dt0[:, None] * geo                     # (n, K)
        x = o[None, None, :] + s[..., None] * d[idx][:, None, :]
        with torch.no_grad():
            f = O.decoder_forward(prior, torch.cat([lat.expand(n * K, -1), x.reshape(-1, 3)], 1)).view(n, K)
        front = tau[idx].clone(); alive = torch.ones(n, dtype=torch.bool); isnear = torch.zeros(n, dtype=torch.bool)
        tnear = torch.zeros(n); lastf = fh[idx].clone(); lasts = tau[idx].clone(); prevf = lastf.clone(); prevs = lasts.clone() - 1
        for k in range(K):
            fk = f[:, k]; sk = s[:, k]
            reach = alive & ((sk - fk.abs() <= front) | (k == 0)) & (sk <= t1[idx])
            nr = reach & (fk.abs() < near)
            isnear |= nr; tnear = torch.where(nr, sk, tnear)
            ok = reach & ~nr & (fk > 0)
            prevf = torch.where(ok, lastf, prevf); prevs = torch.where(ok, lasts, prevs)
            front = torch.where(ok, torch.maximum(front, sk + fk), front)
            lastf = torch.where(ok, fk, lastf); lasts = torch.where(ok, sk, lasts)
            alive = ok
        # new state
        m_new = ((lastf - prevf) / (lasts - prevs).clamp(min=1e-6)).clamp(-1.0, 0.0)
        m_new = torch.where(prevs < lasts, m_new, mh[idx])
        # predicted f at the new front
        f_new = (lastf + m_new * (front - lasts)).clamp(min=0.0)
        f_new = torch.maximum(f_new, 0.25 * (front - lasts))        # the front is at distance >= 0 ... keep a floor
        tau[idx] = torch.where(isnear, tnear, front)
        fh[idx] = f_new; mh[idx] = m_new
        nearset[idx[isnear]] = True
        out = ~isnear & (front > t1[idx])
        active[idx[isnear | out]] = False
        log.append((n, K))
    return log, nearset, tau

# grid start also returns the interpolated value and slope along the ray
def grid_start2(delta=0.03, stop=0.06, max_steps=64):
    active = t0 <= t1
    tau = t0.clone(); near = torch.zeros_like(active); fv = torch.zeros_like(tau); mv = torch.full_like(tau, -1.0)
    for s_ in range(max_steps):
        idx = active.nonzero().squeeze(1)
        if idx.numel() == 0: break
        x = o + tau[idx, None] * d[idx]
        f = interp(x)
        f2 = interp(x + 0.01 * d[idx])
        nr = f < stop
        near[idx[nr]] = True; active[idx[nr]] = False
        fv[idx[nr]] = f[nr]; mv[idx[nr]] = ((f2 - f) / 0.01)[nr].clamp(-1, 0)
        go = ~nr
        tau[idx[go]] += (f[go] - delta)
        out = tau[idx] > t1[idx]
        active[idx[out & go]] = False
    return tau, near, fv, mv

tau_g, near_g, fv, mv = grid_start2(0.03, 0.06)
for c in (0.9, 1.2, 1.4):
    for budget in (18944, 37888):
        log, ns, tau3 = spec_march(tau_g, near_g, fv.clamp(min=0.02), mv, c=c, rows_budget=budget)
        both = ns & ns_ref
        print(f"spec c={c} budget={budget}: launches {len(log)}, rows {sum(n*k for n,k in log)}, rounds {sum(-(-n*k//18944) for n,k in log)}, near {int(ns.sum())} ref {n_ref} common {int(both.sum())} max tau diff {float((tau3-tau_ref)[both].abs().max()):.4f}")
        print("   ", log)
