"""Trace mode (sphere tracing) on the GPU vs its torch oracle (T9) and vs splat mode (T10, sanity)."""
import numpy as np
import pytest
import torch

from oracle import prior as P
from oracle import scenes
from oracle import sdf_oracle as O
from oracle import trace_oracle as T

pytestmark = pytest.mark.gpu
cuda = torch.device("cuda")


def _setup(stock_prior_path, size):
    from sdflabel_b200.deepsdf.workspace import setup_dsdf
    from sdflabel_b200.renderer.tracer import SphereTracer
    dec, L = setup_dsdf(stock_prior_path, precision=torch.float32)
    dec = dec.to(cuda)
    prior = P.load_prior(stock_prior_path)
    K = scenes.intrinsics(size)
    tracer = SphereTracer(K, (size, size)).to(cuda)
    return dec, prior, K, tracer


@pytest.mark.parametrize("size", [48, 96])
def test_trace_maps_and_gradients_vs_oracle(stock_prior_path, size):
    dec, prior, K, tracer = _setup(stock_prior_path, size)
    lat_raw = torch.tensor([0.5, 0.7, 0.5])
    pose0 = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.05, -0.02, 5.0]))
    # oracle
    lat_o = torch.nn.functional.normalize(lat_raw, dim=0).requires_grad_(True)
    pose_o = pose0.clone().requires_grad_(True)
    # (the plain march of the specification needs up to ~250 steps along a grazing ray; the fused march covers up to
    #  32 of them per launch, so the specification gets the steps it needs: same hit set)
    ro = T.trace(prior, lat_o, K, size, size, pose_o, max_steps=256)
    gen = torch.Generator().manual_seed(0)
    cd, cn = torch.rand(ro["depth"].shape, generator=gen), torch.rand(ro["nocs"].shape, generator=gen)
    # ours (latent passed already normalised so the two gradients are of the same variable)
    lat_g = torch.nn.functional.normalize(lat_raw, dim=0).to(cuda).requires_grad_(True)
    pose_g = pose0.clone().to(cuda).requires_grad_(True)
    r = tracer(dec, lat_g, pose_g, normalize_latent=False)
    both = (r["mask"][0].cpu() > 0.5) & (ro["mask"][0] > 0.5)
    either = (r["mask"][0].cpu() > 0.5) | (ro["mask"][0] > 0.5)
    # hit sets agree up to borderline rays: a grazing ray that plain sphere tracing abandons after max_steps may still be
    # resolved by the full-precision finish of the fused march (and the other way round); every such pixel must lie
    # on the silhouette (a 4-neighbour of a miss)
    diff = either & ~both
    assert diff.sum() <= max(4, 0.01 * either.sum()), (int(both.sum()), int(either.sum()))
    miss = torch.nn.functional.pad(~(ro["mask"][0] > 0.5), (1, 1, 1, 1), value=True)
    edge = miss[:-2, 1:-1] | miss[2:, 1:-1] | miss[1:-1, :-2] | miss[1:-1, 2:] | miss[1:-1, 1:-1]
    assert bool(edge[diff].all())
    # the ray parameter of a hit is defined up to the stopping band |f| < eps, i.e. eps / |grad f . d| in tau
    # (oracle/trace_oracle.py); the fused march polishes the root, the oracle stops at the band's outer edge
    band = 1.75 * 1e-4 / ro["slope"][0].detach().clamp(min=1e-3) + 2e-5
    d_err = (r["depth"][0].cpu() - ro["depth"][0].detach()).abs()
    assert bool((d_err[both] <= band[both]).all()), float((d_err / band)[both].max())
    c_err = (r["color"].cpu() - ro["nocs"].detach()).abs().max(dim=0)[0]
    assert bool((c_err[both] <= 0.5 * band[both] + 1e-6).all()), float((c_err / band)[both].max())
    assert float(d_err[both].median()) < 3e-4 and float(c_err[both].median()) < 1.5e-4     # the oracle stops at the band's edge
    # our hits are roots by the specification's own measure: |f| < eps at OUR points, evaluated by the oracle decoder;
    # and the normals are the oracle's gradient at those points (the field is piecewise linear with pieces far
    # smaller than the stopping band, so normals are only comparable at the same point)
    o_, d_, rn_ = T.rays(K, size, size, pose0)
    idx = both.reshape(-1).nonzero().squeeze(1)
    tau_ours = r["depth"][0].detach().cpu().reshape(-1)[idx] / rn_[idx, 2]
    xg = (o_ + tau_ours[:, None] * d_[idx]).requires_grad_(True)
    f_at = O.decoder_forward(prior, torch.cat([lat_o.detach().expand(idx.numel(), -1), xg], 1)).squeeze(1)
    assert float(f_at.abs().max()) < 1e-4 + 2e-5, float(f_at.abs().max())
    (G_at,) = torch.autograd.grad(f_at.sum(), xg)
    n_want = ((G_at / G_at.norm(dim=1, keepdim=True)) @ pose0[:3, :3].t() + 1) / 2
    n_err = (r["normals"].cpu().reshape(3, -1)[:, idx].t() - n_want).abs()
    assert (n_err > 1e-3).float().mean() < 5e-3, float(n_err.max())
    # gradients: the specification's implicit derivative evaluated at OUR hit points (same reason as the normals), with
    # the cotangents restricted to the common hit set so that borderline rays do not enter
    w = both.float()
    tau_all = torch.zeros(size * size)
    tau_all[idx] = tau_ours
    ro2 = T.render_hits(prior, lat_o, K, size, size, pose_o, tau_all, idx)
    lo2 = (ro2["depth"] * cd * w).sum() + (ro2["nocs"] * cn * w).sum()
    g_lat_o, g_pose_o = torch.autograd.grad(lo2, [lat_o, pose_o])
    lg = (r["depth"] * (cd * w).to(cuda)).sum() + (r["color"] * (cn * w).to(cuda)).sum()
    g_lat, g_pose = torch.autograd.grad(lg, [lat_g, pose_g])
    assert np.abs(g_lat.cpu().numpy() - g_lat_o.numpy()).max() < 2e-3 * np.abs(g_lat_o.numpy()).max()
    gp, gpo = g_pose.cpu().numpy()[:3], g_pose_o.numpy()[:3]
    assert np.abs(gp - gpo).max() < 2e-3 * np.abs(gpo).max(), (gp, gpo)


@pytest.mark.parametrize("view", [(0.6, (0.0, 0.0, 5.0), 128), (2.3, (0.2, -0.1, 3.6), 160), (-1.1, (-0.3, 0.15, 7.0), 64)])
def test_trace_fused_march_matches_plain_march(stock_prior_path, view):
    """The fused form (distance cache, speculative look-ahead samples, Newton finish) finds the hits of the plain
    full-precision march (one decoder evaluation per step, tau += sdf; SDFR_MLP_FFMA selects it): same silhouette up to
    rays on its edge, same ray parameter up to the stopping band |f| < eps."""
    from sdflabel_b200 import _lib
    from sdflabel_b200.deepsdf.workspace import setup_dsdf
    yaw, trans, size = view
    dec, prior, K, tracer = _setup(stock_prior_path, size)
    plain, _ = setup_dsdf(stock_prior_path, precision=torch.float32)
    plain = plain.to(cuda)
    plain.mlp_impl = _lib.MLP_FFMA
    from sdflabel_b200.renderer.tracer import SphereTracer
    # the plain march gets the steps a grazing ray needs (the fused one covers up to 32 of them per launch)
    plain_tracer = SphereTracer(K, (size, size), max_steps=256).to(cuda)
    lat = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0).to(cuda)
    pose = O.yaw_pose(torch.tensor([yaw]), torch.tensor(trans)).to(cuda)
    with torch.no_grad():
        a = tracer(dec, lat, pose, normalize_latent=False)
        b = plain_tracer(plain, lat, pose, normalize_latent=False)
    ma, mb = a["mask"][0] > 0.5, b["mask"][0] > 0.5
    both, either = ma & mb, ma | mb
    assert int(both.sum()) > 100
    diff = either & ~both
    assert int(diff.sum()) <= max(4, 0.01 * int(either.sum())), (int(both.sum()), int(either.sum()))
    miss = torch.nn.functional.pad(~mb, (1, 1, 1, 1), value=True)
    edge = miss[:-2, 1:-1] | miss[2:, 1:-1] | miss[1:-1, :-2] | miss[1:-1, 2:] | miss[1:-1, 1:-1]
    assert bool(edge[diff].all())
    # the plain march stops anywhere in |f| < eps, the fused one within |f| < eps / 2 of the root: in tau that is
    # 1.5 eps / |grad f . d|, with the slope the specification's decoder has at the fused hit points
    o_, d_, rn_ = T.rays(K, size, size, pose.cpu())
    idx = both.reshape(-1).nonzero().squeeze(1).cpu()
    tau_all = torch.zeros(size * size)
    tau_all[idx] = a["depth"][0].reshape(-1).cpu()[idx] / rn_[idx, 2]
    slope = T.render_hits(prior, lat.cpu(), K, size, size, pose.cpu(), tau_all, idx)["slope"][0].detach()
    d_err = (a["depth"][0] - b["depth"][0]).abs().cpu()
    band = 1.75 * 1e-4 / slope.clamp(min=1e-3) + 3e-5
    # where the ray grazes the surface (|grad f . d| < 0.05) "the" hit is ambiguous: the plain march may stop at a
    # near-miss minimum with |f| < eps that the fused one passes on its way to the sign change behind it
    steep = both.cpu() & (slope > 0.05)
    assert int(steep.sum()) >= 0.97 * int(both.sum())
    assert bool((d_err[steep] <= band[steep]).all()), float((d_err / band)[steep].max())


def test_trace_agrees_with_splat_mode(stock_prior_path):
    """Sanity (not parity): both renderers see the same surface (SURVEY.md T10 bands)."""
    from sdflabel_b200.grid import Grid3D
    from sdflabel_b200.renderer.rasterer import Rasterer
    size = 96
    dec, prior, K, tracer = _setup(stock_prior_path, size)
    lat = torch.nn.functional.normalize(torch.tensor([0.5, 0.7, 0.5]), dim=0).to(cuda)
    pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0])).to(cuda)
    r = tracer(dec, lat, pose, normalize_latent=False)
    grid = Grid3D(40, device=cuda)
    sdf, _ = dec(torch.cat([lat.expand(grid.points.size(0), -1), grid.points], 1))
    pts, _, nrm = grid.get_surface_points(sdf)
    ras = Rasterer(K, (size, size)).to(cuda)
    rendering, _ = ras(pts, nrm, nrm, pose, rot='dcm', output_mask=True, output_depth=True, output_normals=True,
                       output_nocs=True)
    a, b = r["mask"][0] > 0.5, rendering["mask"][0] > 0.5
    iou = float((a & b).sum()) / float((a | b).sum())
    assert iou > 0.95, iou
    m = a & b
    assert float((r["depth"][0][m] - rendering["depth"][0][m]).abs().median()) < 1e-2
    assert float((r["normals"][:, m] - rendering["normals"][:, m]).abs().median()) < 3e-2


def test_trace_distance_cache_reuse(stock_prior_path):
    """The distance cache of the fused march is kept across calls: another pose of the same latent, and a latent
    moved by less than the Lipschitz slack (lip |dz| <= 0.01, added to the march's margin), must render what a
    tracer that rebuilds the cache every call renders - and a latent moved further must renew the cache."""
    from sdflabel_b200.renderer.tracer import SphereTracer
    size = 96
    dec, prior, K, tracer = _setup(stock_prior_path, size)
    fresh = SphereTracer(K, (size, size), reuse_cache=False).to(cuda)
    lip = float(dec.native().latent_lipschitz)
    assert lip > 0
    z0 = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0)
    step = torch.tensor([1.0, -1.0, 0.0]) / np.sqrt(2.0)
    cases = [(z0, 0.6), (z0, 2.0),                                   # same latent, two poses: reuse with no extra margin
             (z0 + step * (0.006 / lip), 2.0),                       # inside the slack: reuse with a wider margin
             (torch.nn.functional.normalize(z0 + step * 0.05, dim=0), 1.2)]   # far: renewed
    state = lambda: tracer._dist_cache[1][-256:-240].view(torch.int32).cpu().tolist()
    rows = []
    for z, yaw in cases:
        pose = O.yaw_pose(torch.tensor([yaw]), torch.tensor([0.0, 0.0, 5.0])).to(cuda)
        with torch.no_grad():
            a = tracer(dec, z.to(cuda), pose, normalize_latent=False)
            b = fresh(dec, z.to(cuda), pose, normalize_latent=False)
        rows.append(state()[1])
        ma, mb = a["mask"][0] > 0.5, b["mask"][0] > 0.5
        both, either = ma & mb, ma | mb
        assert int(both.sum()) > 100
        assert int((either & ~both).sum()) <= max(2, 0.005 * int(either.sum()))
        # both stop within |f| < eps / 2 of the same root
        o_, d_, rn_ = T.rays(K, size, size, pose.cpu())
        idx = both.reshape(-1).nonzero().squeeze(1).cpu()
        tau_all = torch.zeros(size * size)
        tau_all[idx] = a["depth"][0].reshape(-1).cpu()[idx] / rn_[idx, 2]
        slope = T.render_hits(prior, z, K, size, size, pose.cpu(), tau_all, idx)["slope"][0].detach()
        steep = both.cpu() & (slope > 0.05)
        d_err = (a["depth"][0] - b["depth"][0]).abs().cpu()
        band = 1.25 * 1e-4 / slope.clamp(min=1e-3) + 3e-5
        assert bool((d_err[steep] <= band[steep]).all()), float((d_err / band)[steep].max())
    assert rows[0] == 40 ** 3 and rows[1] == 0 and rows[2] == 0 and rows[3] == 40 ** 3, rows


def test_trace_views_in_flight_match_sequential(stock_prior_path):
    """render_views (several poses of one latent on concurrent CUDA streams, one shared distance cache) returns
    what one forward per pose returns: the rays of a view do not interact with anything else in flight."""
    size = 64
    dec, prior, K, tracer = _setup(stock_prior_path, size)
    lat = torch.tensor([0.6, 0.6, 0.5], device=cuda)
    poses = [O.yaw_pose(torch.tensor([0.3 + 0.9 * i]), torch.tensor([0.05 * i, 0.0, 4.0 + 0.5 * i])) for i in range(6)]
    for _ in range(2):                                    # second pass: every stream reuses its cache
        outs = tracer.render_views(dec, lat, poses, views_in_flight=3)
        assert len(outs) == len(poses)
        for pose, got in zip(poses, outs):
            with torch.no_grad():
                want = tracer(dec, lat, pose.to(cuda))
            assert int(want["mask"].sum()) > 50
            assert torch.equal(got["mask"], want["mask"])
            for key in ("depth", "color", "normals"):
                assert float((got[key] - want[key]).abs().max()) <= 1e-6, key


def test_trace_empty_view(stock_prior_path):
    """Camera looking away from the object: no hits, zero maps, zero gradients."""
    dec, prior, K, tracer = _setup(stock_prior_path, 32)
    lat = torch.tensor([0.5, 0.7, 0.5], device=cuda, requires_grad=True)
    pose = O.yaw_pose(torch.tensor([0.0]), torch.tensor([30.0, 0.0, 5.0])).to(cuda).requires_grad_(True)
    r = tracer(dec, lat, pose)
    assert float(r["mask"].sum()) == 0.0 and float(r["depth"].abs().sum()) == 0.0
    g = torch.autograd.grad(r["depth"].sum() + r["color"].sum(), [lat, pose], allow_unused=True)
    assert all(x is None or float(x.abs().sum()) == 0.0 for x in g)


def _octahedron_decoder(radius, latent_size=3):
    """A Decoder that IS a closed-form field: f(l, x) = tanh((|x|_1 - r) / sqrt(3)), the exact distance to the
    faces of the octahedron |x|_1 <= r (a lower bound of the distance near its edges), as one hidden ReLU layer:
    |x_i| = relu(x_i) + relu(-x_i).  The latent columns carry zero weights."""
    from sdflabel_b200.deepsdf.networks.deep_sdf_decoder_scale import Decoder
    dec = Decoder(latent_size, [8], dropout=None, dropout_prob=0.0, norm_layers=(), latent_in=(), weight_norm=False)
    w0 = torch.zeros(8, latent_size + 3)
    for a in range(3):
        w0[2 * a, latent_size + a] = 1.0
        w0[2 * a + 1, latent_size + a] = -1.0
    w1 = torch.zeros(1, 8)
    w1[0, :6] = 1.0 / np.sqrt(3.0)
    sd = dec.state_dict()
    sd["lin0.weight"], sd["lin0.bias"] = w0, torch.zeros(8)
    sd["lin1.weight"], sd["lin1.bias"] = w1, torch.tensor([-radius / np.sqrt(3.0)])
    dec.load_state_dict(sd)
    return dec.to(cuda).eval()


@pytest.mark.parametrize("impl", ["auto", "ffma"])
def test_trace_closed_form_octahedron(impl):
    """T9, closed form: the traced depth, NOCS and normals of an analytic field against the exact ray / octahedron
    intersection (eight planes).  Sphere tracing stops at |f| < eps, i.e. within eps / cos(incidence) of the face."""
    from sdflabel_b200 import _lib
    from sdflabel_b200.renderer.tracer import SphereTracer
    size, radius, eps = 64, 0.6, 1e-4
    dec = _octahedron_decoder(radius)
    if impl == "ffma":
        dec.mlp_impl = _lib.MLP_FFMA
    K = scenes.intrinsics(size)
    pose = O.yaw_pose(torch.tensor([0.4]), torch.tensor([0.03, -0.05, 4.0]))
    tracer = SphereTracer(K, (size, size), eps=eps).to(cuda)
    r = tracer(dec, torch.tensor([0.5, 0.7, 0.5], device=cuda), pose.to(cuda))
    # exact intersection in float64
    o, d, rn = (t.double() for t in T.rays(K.double(), size, size, pose.double()))
    signs = torch.tensor([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=torch.float64)
    sd_, so_ = d @ signs.t(), (signs @ o)[None, :]                     # (P,8) s.d, (1,8) s.o
    tt = (radius - so_) / sd_
    t_in = torch.where(sd_ < 0, tt, torch.full_like(tt, -1e30)).max(dim=1)
    t_out = torch.where(sd_ > 0, tt, torch.full_like(tt, 1e30)).min(dim=1)[0]
    hit = (t_in[0] <= t_out) & (t_out > 0)
    face = signs[t_in[1]] / np.sqrt(3.0)                               # outward normal of the entry face
    cosi = (face * d).sum(1).abs()
    got_hit = r["mask"][0].cpu().reshape(-1) > 0.5
    # rays that pass the octahedron closer than the stopping band may legitimately count as hits
    both = got_hit & hit
    assert both.sum() >= 0.99 * (got_hit | hit).sum() and both.sum() > 200, (int(both.sum()), int(hit.sum()))
    tau = (r["depth"][0].cpu().reshape(-1).double() / rn[:, 2])[both]
    err = (tau - t_in[0][both]).abs()
    bound = 2.0 * eps / cosi[both] + 2e-5
    assert bool((err <= bound).all()), (float(err.max()), float((err / bound).max()))
    x = o + t_in[0][both, None] * d[both]
    interior = x.abs().min(dim=1)[0] > 5e-3                            # away from the edges, where the field has kinks
    nocs = r["color"].cpu().reshape(3, -1).double()[:, both]
    want = ((x * torch.tensor([-1.0, 1.0, 1.0], dtype=torch.float64) + 1) / 2).t()
    assert float(((nocs - want).abs().max(dim=0)[0] - bound / 2).max()) <= 0.0
    n_cam = face[both] @ pose[:3, :3].double().t()
    got_n = r["normals"].cpu().reshape(3, -1).double()[:, both]
    assert float(((got_n - ((n_cam + 1) / 2).t()).abs().max(dim=0)[0])[interior].max()) < 1e-5
