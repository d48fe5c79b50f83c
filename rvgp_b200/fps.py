"""K1 host wrapper: furthest-point sampling on the device (reference RVGP/geometry.py:126-162)."""
import torch

from ._cabi import get_handle, I64


def furthest_point_sampling_device(Xd, N=None, spacing=0.1, start_idx=0):
    """Xd (n, D) cuda f64.  Returns (perm int32 cuda, lambdas f64 cuda), truncated like the reference."""
    h = get_handle(Xd.device.index)
    n, D = Xd.shape
    cap = n if N is None else int(N)
    perm = torch.zeros(cap, dtype=torch.int32, device=Xd.device)
    lambdas = torch.zeros(cap, dtype=torch.float64, device=Xd.device)
    count = torch.zeros(1, dtype=torch.int32, device=Xd.device)
    wsb = h.query("rvgp_fps_workspace_bytes", h._h, int(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=Xd.device)
    h.call("rvgp_fps_f64", Xd, int(n), int(D), 0 if N is None else int(N), float(spacing), int(start_idx),
           perm, lambdas, count, ws, I64(wsb))
    c = int(count.item())
    return perm[:c], lambdas[:c]
