"""Camera trajectory -> (intrinsics, relative camera-to-world poses): the host half of the Plucker-embedding producer
(SURVEY 8f-3).  Restates ``Camera`` (src/dataset/dance_image_h_v_camera.py:17-77), ``get_relative_pose`` and the intrinsics /
pose assembly of ``camera_file_to_embedding`` (scripts/pose2vid.py:29-84); the per-pixel ray arithmetic of ``ray_condition``
(dance_image_h_v_camera.py:88-130) runs on the GPU inside ``CameraPoseEncoder.forward_cameras`` (hv_op_plucker_unshuffle), so the
(1, 6, F, H, W) embedding -- 127 MB in fp16 at 24x768x576, rebuilt on the CPU and copied per clip by the reference -- never exists.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

# scripts/pose2vid.py:58-61
STATIC_CAMERA_LANDSCAPE = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.788079, 1.0]
STATIC_CAMERA_PORTRAIT = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.788079, 1.0, 1.0]


class Camera:
    """One line ``ts tx ty tz qx qy qz qw fx fy [scale]`` of a camera file (dance_image_h_v_camera.py:17-77)."""

    def __init__(self, entry: Sequence[float], pose_file_name: str = "test", image_scale=(1920, 1080)):
        assert len(entry) in (10, 11), f"length of entry should be 11 (extrinsic + fx fy + scale) or 10 (+ fx fy), got {len(entry)}"
        if image_scale[0] > image_scale[1]:
            self.fx = entry[8]
            self.fy = self.fx * (image_scale[0] / image_scale[1])
        else:
            self.fy = entry[9]
            self.fx = self.fy * (image_scale[1] / image_scale[0])
        self.cx = self.cy = 0.5
        self.timestamp = entry[0]
        t = np.array(entry[1:4], dtype=np.float64)
        q = np.array(entry[4:8], dtype=np.float64)
        q = q / np.linalg.norm(q)
        qx, qy, qz, qw = q
        rot = np.array([[1 - 2 * qy**2 - 2 * qz**2, 2 * qx * qy - 2 * qz * qw, 2 * qx * qz + 2 * qy * qw],
                        [2 * qx * qy + 2 * qz * qw, 1 - 2 * qx**2 - 2 * qz**2, 2 * qy * qz - 2 * qx * qw],
                        [2 * qx * qz - 2 * qy * qw, 2 * qy * qz + 2 * qx * qw, 1 - 2 * qx**2 - 2 * qy**2]])
        scale = entry[10] if len(entry) == 11 else 1.0
        name = pose_file_name
        if any(k in name for k in ("bedlam", "blender", "ue_rendered")):       # the file stores world-to-camera
            self.w2c_mat = np.eye(4)
            self.w2c_mat[:3, :3] = rot
            self.w2c_mat[:3, 3] = t
            self.c2w_mat = np.linalg.inv(self.w2c_mat)
        elif any(k in name for k in ("pexels", "inference", "ubc", "tiktok", "webvid", "test")):   # camera-to-world, translation * scale
            self.c2w_mat = np.eye(4)
            self.c2w_mat[:3, :3] = rot
            self.c2w_mat[:3, 3] = t * scale
            self.w2c_mat = np.linalg.inv(self.c2w_mat)
        else:
            raise ValueError(f"Unknown camera pose dataset name: {pose_file_name}")


def load_cameras(pose_file: str, img_size) -> List[Camera]:
    with open(pose_file) as f:
        rows = [[float(x) for x in line.strip().split(" ")] for line in f if line.strip()]
    return [Camera(r, pose_file, img_size) for r in rows]


def relative_cameras(cam_params: Sequence[Camera], ref_idx: int, tgt_idx: Sequence[int], img_size) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> intrinsics (1, F, 4) = (fx, fy, cx, cy) in pixels and c2w (1, F, 4, 4) relative to the reference camera, both float32:
    exactly what scripts/pose2vid.py:66-80 passes to ``ray_condition`` (get_relative_pose with the reference view at the origin)."""
    cams = [cam_params[ref_idx]] + [cam_params[i] for i in tgt_idx]
    K = np.asarray([[c.fx * img_size[0], c.fy * img_size[1], c.cx * img_size[0], c.cy * img_size[1]] for c in cams[1:]], dtype=np.float32)
    abs2rel = np.eye(4) @ cams[0].w2c_mat
    rel = np.array([np.eye(4)] + [abs2rel @ c.c2w_mat for c in cams[1:]], dtype=np.float32)[1:]
    return torch.as_tensor(K)[None], torch.as_tensor(rel)[None]


def camera_file_to_cameras(video_length: int, camera_file: str, ref_img_idx: int, tgt_img_idx: Sequence[int], img_size):
    """scripts/pose2vid.py:52-84 up to (not including) ray_condition; a missing file means the static camera (:55-62)."""
    if not os.path.exists(camera_file):
        static = STATIC_CAMERA_LANDSCAPE if img_size[0] > img_size[1] else STATIC_CAMERA_PORTRAIT
        cams = [Camera(static, "test", img_size)] * video_length
    else:
        cams = load_cameras(camera_file, img_size)
    return relative_cameras(cams, ref_img_idx, tgt_img_idx, img_size)
