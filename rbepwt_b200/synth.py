"""Synthetic inputs of the shapes BASELINE.json names (host-side numpy; not on the hot path).

scikit-image is not installed in this image, so Felzenszwalb label maps cannot be produced by
the reference's own call (rbepwt.py:780); label maps are an INPUT to the path (north_star) and
are synthesised here as warped Voronoi diagrams whose cell count / mean region size mimic
felzenszwalb(scale=200, sigma=2, min_size=10) on natural images.
All generators are deterministic in `seed`.
"""
import numpy as np


def voronoi_labels(h, w, n_seeds, seed=0, warp=3.0, shuffle_ids=True):
    """Label map of `n_seeds` connected, ragged-bordered cells.

    Nearest-seed assignment of smoothly displaced pixel coordinates (the displacement field has
    amplitude `warp` px and a long wavelength, so cells stay connected).  Label VALUES are a
    random permutation, so that region order (first appearance, rbepwt.py:840-848) differs from
    label order.
    """
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, [h, w], size=(n_seeds, 2))
    ii, jj = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    ph = rng.uniform(0, 2 * np.pi, size=8)
    k = 2 * np.pi / max(16.0, min(h, w) / 6.0)
    di = warp * (np.sin(k * jj + ph[0]) + 0.6 * np.sin(2.3 * k * ii + ph[1]) + 0.4 * np.sin(3.1 * k * (ii + jj) + ph[2]))
    dj = warp * (np.sin(k * ii + ph[3]) + 0.6 * np.sin(2.7 * k * jj + ph[4]) + 0.4 * np.sin(3.7 * k * (ii - jj) + ph[5]))
    q = np.stack([(ii + di).ravel(), (jj + dj).ravel()], axis=1)
    _, idx = cKDTree(pts).query(q)
    lab = idx.reshape(h, w).astype(np.int32)
    if shuffle_ids:
        lab = rng.permutation(n_seeds).astype(np.int32)[lab]
    return np.ascontiguousarray(lab)


def heavytail_labels(n, n_seeds, seed=0):
    """n x n label map with a heavy-tailed region-size distribution -- a few 10^4-pixel background regions next
    to hundreds of small ones, which is what felzenszwalb(scale=200) produces on natural images: 85 % of the
    Voronoi seeds crowd into three blobs covering ~15 % of the image, the rest are spread out."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.15 * n, 0.85 * n, size=(3, 2))
    ncrowd = int(0.85 * n_seeds)
    crowd = centres[rng.integers(0, 3, size=ncrowd)] + rng.normal(0, 0.07 * n, size=(ncrowd, 2))
    spread = rng.uniform(0, n, size=(n_seeds - ncrowd, 2))
    pts = np.clip(np.concatenate([crowd, spread]), 0, n - 1)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    _, idx = cKDTree(pts).query(np.stack([ii.ravel(), jj.ravel()], 1))
    return np.ascontiguousarray(idx.reshape(n, n).astype(np.int32))


def block_labels(h, w, b):
    """Grid of b x b blocks (the label map BASELINE.md section 2 timed the reference on)."""
    ii, jj = np.meshgrid(np.arange(h) // b, np.arange(w) // b, indexing="ij")
    return np.ascontiguousarray((ii * ((w + b - 1) // b) + jj).astype(np.int32))


def piecewise_smooth_image(labels, seed=0, noise=2.0):
    """float64 image: one affine ramp per region + Gaussian noise, clipped to [0,255].
    Continuous-valued, so top-k and EPWT ties have probability ~0."""
    rng = np.random.default_rng(seed)
    h, w = labels.shape
    _, inv = np.unique(labels, return_inverse=True)
    inv = inv.reshape(h, w)
    nreg = int(inv.max()) + 1
    base = rng.uniform(20, 235, size=nreg)
    sl = rng.uniform(-0.5, 0.5, size=(nreg, 2))
    ii, jj = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    img = base[inv] + sl[inv, 0] * (ii - h / 2) + sl[inv, 1] * (jj - w / 2)
    img = img + rng.normal(0, noise, size=(h, w))
    return np.ascontiguousarray(np.clip(img, 0.0, 255.0))


def smooth_field_image(h, w, seed=0, sigma=8.0):
    """float64 Gaussian-filtered noise scaled to [0,255] (the EPWT config's image)."""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(seed)
    f = gaussian_filter(rng.normal(size=(h, w)), sigma, mode="wrap")
    f = (f - f.min()) / (f.max() - f.min())
    return np.ascontiguousarray(255.0 * f)


def noise_image(h, w, seed=0):
    return np.ascontiguousarray(np.random.default_rng(seed).uniform(0, 255, size=(h, w)))


def config_inputs(name, seed=None):
    """(img, labels, kwargs) for the named BASELINE.json config."""
    if name == "cameraman256":
        s = 0 if seed is None else seed
        lab = voronoi_labels(256, 256, 384, seed=s)
        return piecewise_smooth_image(lab, seed=s), lab
    if name == "synthetic512":
        s = 1 if seed is None else seed
        lab = voronoi_labels(512, 512, 1024, seed=s)
        return piecewise_smooth_image(lab, seed=s), lab
    if name == "epwt512":
        s = 2 if seed is None else seed
        return smooth_field_image(512, 512, seed=s), None
    if name == "small2048":
        s = 3 if seed is None else seed
        lab = voronoi_labels(2048, 2048, 65536, seed=s, warp=1.0)
        return piecewise_smooth_image(lab, seed=s), lab
    raise KeyError(name)


def torch_batch(nimg, h, w, n_seeds, seed0, device="cpu", noise=2.0, warp=3.0):
    """`nimg` distinct (image float64 [h,w], labels int32 [h,w]) pairs as torch tensors on `device`:
    the config-2/5 generator (warped Voronoi labels + per-region affine ramps + noise), image i seeded
    with seed0 + i.  All randomness is drawn on the host with numpy, so CPU and CUDA runs see the same
    inputs; only the nearest-seed search runs on `device` (benchmark plumbing, not the hot path)."""
    import torch

    dev = torch.device(device)
    imgs = torch.empty((nimg, h, w), dtype=torch.float64, device=dev)
    labs = torch.empty((nimg, h, w), dtype=torch.int32, device=dev)
    ii, jj = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev),
                            torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    k = 2 * np.pi / max(16.0, min(h, w) / 6.0)
    rows = max(1, (1 << 26) // max(n_seeds, 1))
    for i in range(nimg):
        rng = np.random.default_rng(seed0 + i)
        pts = torch.from_numpy(rng.uniform(0, [h, w], size=(n_seeds, 2)).astype(np.float32)).to(dev)
        ph = rng.uniform(0, 2 * np.pi, size=6)
        ids = torch.from_numpy(rng.permutation(n_seeds).astype(np.int32)).to(dev)
        base = torch.from_numpy(rng.uniform(20, 235, size=n_seeds)).to(dev)
        sl = torch.from_numpy(rng.uniform(-0.5, 0.5, size=(n_seeds, 2))).to(dev)
        nz = torch.from_numpy(rng.normal(0, noise, size=(h, w))).to(dev)
        di = warp * (torch.sin(k * jj + ph[0]) + 0.6 * torch.sin(2.3 * k * ii + ph[1])
                     + 0.4 * torch.sin(3.1 * k * (ii + jj) + ph[2]))
        dj = warp * (torch.sin(k * ii + ph[3]) + 0.6 * torch.sin(2.7 * k * jj + ph[4])
                     + 0.4 * torch.sin(3.7 * k * (ii - jj) + ph[5]))
        q = torch.stack([(ii + di).reshape(-1), (jj + dj).reshape(-1)], dim=1)
        cell = torch.empty(h * w, dtype=torch.int64, device=dev)
        p2 = (pts * pts).sum(1)
        for a in range(0, h * w, rows):
            qa = q[a:a + rows]
            cell[a:a + rows] = (p2[None, :] - 2.0 * (qa @ pts.T)).argmin(dim=1)
        cell = cell.reshape(h, w)
        labs[i] = ids[cell]
        x = base[cell] + sl[cell, 0] * (ii.double() - h / 2) + sl[cell, 1] * (jj.double() - w / 2) + nz
        imgs[i] = x.clamp_(0.0, 255.0)
    return imgs, labs
