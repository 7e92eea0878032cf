"""TEST INFRASTRUCTURE ONLY (imported by tests/ -- never by the product path).

CPU restatement of the keyframe hand-over at the end of the reference's frontend, the oracle of
`super_primitive_b200.handover` (compaction straight from `integrated_depth`):

    frontend/process_frame.py:231-236   nearest resampling of integrated_depth to the keyframe grid, masks = depth > 1e-7,
                                        put_keypoints_back, logdepth[masks] = log(logdepth[masks])
    image/keyframe.py:151-173           put_keypoints_back: drop empty segments, move every keypoint to the nearest mask
                                        pixel (Euclidean distance to its rounded pixel position, argmin = first among equals)
    tool/point_utils.py:31-40           (de)normalisation of the keypoints

Pinned: tests/golden/make_golden_handover.py runs the reference's own `put_keypoints_back` (imported from the checkout)
inside the four lines of `process_to_kf` and asserts this module reproduces every output bit for bit (handover.npz).
"""
import torch


def denormalise(x_norm, dims):
    dims = torch.as_tensor(dims, dtype=torch.float32)
    return (0.5 * (dims - 1) * (x_norm + 1)).round().long()


def normalise(x_pixel, dims):
    inv = 1.0 / (torch.as_tensor(dims, dtype=torch.float32) - 1)
    return 2 * x_pixel * inv - 1


def snap_keypoints(keypoints, masks, logdepth):
    """image/keyframe.py:151-173"""
    _, H, W = masks.shape
    kp = denormalise(keypoints, (H, W))
    good = masks.sum(dim=(1, 2)) > 0
    kp, masks, logdepth = kp[good], masks[good], logdepth[good]
    for i in range(kp.shape[0]):
        r, c = kp[i]
        rows, cols = torch.where(masks[i])
        d = torch.sqrt((rows - r) ** 2 + (cols - c) ** 2)
        j = torch.argmin(d)
        kp[i] = torch.stack([rows[j], cols[j]])
    return normalise(kp, (H, W)), masks, logdepth, good


def handover(integrated_depth, keypoints, size):
    """frontend/process_frame.py:231-236.  Returns (keypoints (M,2), masks (M,H,W), logdepth (M,H,W), good (N,))."""
    logdepth = torch.nn.functional.interpolate(integrated_depth[:, None], size=tuple(size), mode='nearest')[:, 0]
    masks = logdepth > 1e-7
    keypoints, masks, logdepth, good = snap_keypoints(keypoints, masks, logdepth)
    logdepth = logdepth.clone()
    logdepth[masks] = torch.log(logdepth[masks])
    return keypoints, masks, logdepth, good
