"""``distCUDA2(points)``: mean squared distance of every point to its 3 nearest neighbours
(submodules/simple-knn/spatial.cu + simple_knn.cu:185-220), backed by sgb_knn_mean_dist2 (csrc/knn.cu)."""
import torch

from .. import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """points (P,3) float32 CUDA -> (P,) float32 CUDA.  Same checks as the reference's pybind glue would
    need: a CUDA float tensor of shape (P,3); anything else raises instead of reading garbage."""
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise ValueError("distCUDA2 expects a CUDA tensor (there is no CPU path)")
    if points.ndim != 2 or points.shape[1] != 3:
        raise ValueError("points must have shape (P, 3)")
    pts = points.detach().to(torch.float32).contiguous()
    P = pts.shape[0]
    out = torch.zeros((P,), dtype=torch.float32, device=pts.device)   # spatial.cu: torch::full({P}, 0.0)
    if P == 0:
        return out
    with torch.cuda.device(pts.device):
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        ctx = _lib.ctx_for(pts.device.index, stream)
        _lib.check(_lib.load().sgb_knn_mean_dist2(ctx, P, pts.data_ptr(), out.data_ptr(), stream), "sgb_knn_mean_dist2")
    return out
