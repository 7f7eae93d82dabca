"""Driver-side pieces of the reference's render script, around the fused renderer:

* `build_render_imgs_info`   utils/imgs_info.py:158-181   query pose -> que_imgs_info (all ERP pixels, c2w / w2c, depth range)
* `imgs_info_to_torch`, `to_cuda`                          utils/imgs_info.py (dict plumbing)
* `render_poses`             render.py:249-291            the pose loop: one `renderer(data)` per query pose, images + depths back
* `color_map_backward`, `depth_to_uint8`                   utils/base_utils.py, render.py:68-84 (what save_renderings / save_depth write)
* `WSPSNR`                   network/metrics.py:118-170   latitude-weighted PSNR of equirectangular images

No file IO and no dataset readers here: the reference's LMDB / habitat readers feed the same dictionaries this module takes.
"""
import numpy as np
import torch


def build_render_imgs_info(que_pose, que_shape, que_depth_range):
    """utils/imgs_info.py:158-181.  que_pose (3,4) world->camera, que_shape (h,w), que_depth_range (2,) -> dict of numpy arrays."""
    h, w = int(que_shape[0]), int(que_shape[1])
    que_coords = np.stack(np.meshgrid(np.arange(w), np.arange(h)), -1).reshape([1, -1, 2]).astype(np.float32)
    que_pose = np.asarray(que_pose)
    c2w = np.linalg.inv(np.concatenate([que_pose, np.array([[0, 0, 0, 1]])], axis=0))[:3, :4]
    return {"w2c": que_pose.astype(np.float32)[np.newaxis, :, :], "c2w": c2w.astype(np.float32)[np.newaxis, ...],
            "rots": que_pose[:3, :3].astype(np.float32)[np.newaxis, ...], "trans": que_pose[:3, 3].astype(np.float32)[np.newaxis, ...],
            "coords": que_coords, "depth_range": np.asarray(que_depth_range, np.float32)[None, :], "shape": (h, w)}


def imgs_info_to_torch(imgs_info):
    """numpy leaves -> torch tensors (images (n,h,w,3) uint8-range arrays are expected already as float (n,3,h,w), as the reference's
    build_imgs_info produces them)."""
    return {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in imgs_info.items()}


def to_cuda(data, device="cuda"):
    if isinstance(data, torch.Tensor):
        return data.to(device)
    if isinstance(data, dict):
        return {k: to_cuda(v, device) for k, v in data.items()}
    if isinstance(data, (list, tuple)) and data and isinstance(data[0], torch.Tensor):
        return type(data)(to_cuda(v, device) for v in data)
    return data


def color_map_backward(rgb):
    """utils/base_utils.py: float colours in [0,1] -> uint8."""
    rgb = np.asarray(rgb) * 255
    return np.clip(rgb, a_min=0, a_max=255).astype(np.uint8)


def depth_to_uint8(depth, near, far):
    """render.py:68-84 (save_depth): clip to the range, normalised inverse depth, 8 bits."""
    d = np.clip(np.asarray(depth, dtype=np.float32), a_min=near, a_max=far)
    d = (1 / d - 1 / near) / (1 / far - 1 / near)
    return np.uint8(d * 255)


@torch.no_grad()
def render_poses(renderer, ref_imgs_info, que_poses, que_shapes, que_depth_ranges, src_imgs_info=None, device="cuda",
                 render_depth=True, on_view=None):
    """render.py:249-291 without the file IO: for every query pose build que_imgs_info, call `renderer(data)` (the reference's
    `data = {'que_imgs_info', 'ref_imgs_info', 'src_imgs_info', 'eval'}`), and collect the fine image (uint8, h,w,3) and the
    normalised fine depth (uint8, h,w).  `on_view(qi, render_info)` may consume the raw outputs (e.g. to save them)."""
    ref_d = to_cuda(ref_imgs_info, device)
    src_d = to_cuda(src_imgs_info, device) if src_imgs_info is not None else None
    imgs, depths = [], []
    for qi in range(len(que_poses)):
        que = to_cuda(imgs_info_to_torch(build_render_imgs_info(que_poses[qi], que_shapes[qi], que_depth_ranges[qi])), device)
        data = {"que_imgs_info": que, "ref_imgs_info": dict(ref_d), "eval": True}
        if src_d is not None:
            data["src_imgs_info"] = dict(src_d)
        info = renderer(data)
        h, w = int(que_shapes[qi][0]), int(que_shapes[qi][1])
        key = "pixel_colors_nr_fine" if "pixel_colors_nr_fine" in info else "pixel_colors_nr"
        imgs.append(color_map_backward(info[key].float().cpu().numpy().reshape(h, w, 3)))
        dkey = "render_depth_fine" if "render_depth_fine" in info else "render_depth"
        if render_depth and dkey in info:
            near, far = float(que_depth_ranges[qi][0]), float(que_depth_ranges[qi][1])
            depths.append(depth_to_uint8(info[dkey].float().cpu().numpy().reshape(h, w), near, far))
        if on_view is not None:
            on_view(qi, info)
    return imgs, depths


class WSPSNR:
    """network/metrics.py:118-170: PSNR with the sin(latitude) weights of the equirectangular projection."""

    def __init__(self):
        self.weight_cache = {}

    def get_weights(self, height=1080, width=1920):
        key = str(height) + ";" + str(width)
        if key not in self.weight_cache:
            v = np.sin((np.arange(0, height) + 0.5) * (np.pi / height)).reshape(height, 1)
            self.weight_cache[key] = np.broadcast_to(v, (height, width)).copy()
        return self.weight_cache[key]

    def calculate_wsmse(self, reconstructed, reference):
        """images as (B,H,W,C) tensors -> (B,) weighted mse"""
        b, h, w, c = reconstructed.shape
        weights = torch.tensor(self.get_weights(h, w), device=reconstructed.device, dtype=reconstructed.dtype)
        weights = weights.view(1, h, w, 1).expand(b, -1, -1, c)
        sq = torch.pow(reconstructed - reference, 2.0)
        return torch.sum(weights * sq, dim=(1, 2, 3)) / torch.sum(weights, dim=(1, 2, 3))

    def ws_psnr(self, y_pred, y_true, max_val=1.0):
        return 10 * torch.log10(max_val * max_val / self.calculate_wsmse(y_pred, y_true))
