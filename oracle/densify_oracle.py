"""TEST INFRASTRUCTURE ONLY -- device-agnostic torch restatement of the reference's adaptive density control
(2d-gaussian-splatting/scene/gaussian_model.py:528-647), written as one function on plain tensors.  Checked on the CPU
against golden vectors produced by the reference's own GaussianModel methods (tests/golden/make_golden_densify.py)."""
import torch


def build_rotation(r):
    # utils/general_utils.py:build_rotation
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


KEYS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def densify_and_prune(params, moments, accum, denom, percent_dense, max_grad, min_opacity, extent, max_screen_size, samples, N=2):
    """params / moments: dict name -> tensor / (exp_avg, exp_avg_sq).  samples: the [N*S,3] normal draw of densify_and_split.
    Returns (new params, new moments) in the reference's final row order."""
    p = {k: v.clone() for k, v in params.items()}
    m = {k: (a.clone(), b.clone()) for k, (a, b) in moments.items()}
    grads = accum / denom                                  # :626-627
    grads[grads.isnan()] = 0.0
    get_scaling = lambda: torch.exp(p["scaling"])

    def cat(new):                                          # densification_postfix / cat_tensors_to_optimizer
        for k in KEYS:
            p[k] = torch.cat((p[k], new[k]), dim=0)
            m[k] = (torch.cat((m[k][0], torch.zeros_like(new[k])), dim=0), torch.cat((m[k][1], torch.zeros_like(new[k])), dim=0))

    def prune(mask):                                       # prune_points / _prune_optimizer
        keep = ~mask
        for k in KEYS:
            p[k] = p[k][keep]
            m[k] = (m[k][0][keep], m[k][1][keep])

    # densify_and_clone :601-617
    sel = torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(get_scaling(), dim=1).values <= percent_dense * extent)
    cat({k: p[k][sel] for k in KEYS})
    # densify_and_split :569-599
    n_init = p["xyz"].shape[0]
    padded = torch.zeros((n_init,), device=grads.device)
    padded[:grads.shape[0]] = grads.squeeze()
    sel = torch.where(padded >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(get_scaling(), dim=1).values > percent_dense * extent)
    rots = build_rotation(p["rotation"][sel]).repeat(N, 1, 1)
    new = {"xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + p["xyz"][sel].repeat(N, 1),
           "scaling": torch.log(get_scaling()[sel].repeat(N, 1) / (0.8 * N)),
           "rotation": p["rotation"][sel].repeat(N, 1), "f_dc": p["f_dc"][sel].repeat(N, 1, 1),
           "f_rest": p["f_rest"][sel].repeat(N, 1, 1), "opacity": p["opacity"][sel].repeat(N, 1)}
    cat(new)
    prune(torch.cat((sel, torch.zeros(N * int(sel.sum()), device=sel.device, dtype=bool))))
    # final prune :632-636 (max_radii2D is all zeros after densification_postfix: the screen-size test is vacuous)
    mask = (torch.sigmoid(p["opacity"]) < min_opacity).squeeze()
    if max_screen_size:
        big_vs = torch.zeros_like(mask)
        big_ws = get_scaling().max(dim=1).values > 0.1 * extent
        mask = torch.logical_or(torch.logical_or(mask, big_vs), big_ws)
    prune(mask)
    return p, m
