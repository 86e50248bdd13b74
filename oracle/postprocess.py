"""CPU oracle for the output post-processing that follows the sample loop.  TEST INFRASTRUCTURE ONLY.

numpy fp32 restatement of
* ``datasets/show.py:157-162`` ``inv_standardize`` as the SHOW trainer applies it (trainers/ddpm_show_trainer.py:913-921),
* the BEAT trainer's axis-angle branch (trainers/ddpm_beat_trainer.py:1056-1062): de-normalise the axis-angle gesture,
  ``axis_angle_to_euler_angles`` (datasets/rotation_converter.py:282-296 = ``matrix_to_euler_angles(axis_angle_to_matrix(.),
  'XYZ')``: rc:204-233 axis-angle -> quaternion, rc:251-280 quaternion -> matrix, rc:342-385 + rc:299-329 matrix -> XYZ Euler),
  degrees, re-normalise with the Euler statistics.
Pinned by tests/test_postprocess.py against the reference functions themselves (fixture tests/golden/postprocess_beat.npz,
made by tests/golden/make_golden.py, and live when /root/reference is present).
"""
import numpy as np

F = np.float32


def inv_standardize(data, mean, std):
    """show.py:159: ``data * std + mean`` (two roundings, no FMA)."""
    return (np.asarray(data, F) * np.asarray(std, F)).astype(F) + np.asarray(mean, F)


def axis_angle_to_euler_xyz(aa):
    """rc:282-296 for ``aa[..., 3]`` (radians) -> XYZ Euler angles (radians), fp32 like the reference."""
    aa = np.asarray(aa, F)
    angles = np.sqrt((aa * aa).sum(-1, keepdims=True, dtype=F)).astype(F)          # rc:217 torch.norm(p=2)
    half = angles * F(0.5)
    small = np.abs(angles) < F(1e-6)                                               # rc:219-220
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(small, F(0.5) - (angles * angles) / F(48), np.sin(half) / angles).astype(F)   # rc:221-229
    r = np.cos(half)[..., 0].astype(F)
    i, j, k = [(aa[..., n] * s[..., 0]).astype(F) for n in range(3)]               # rc:230-232
    two_s = (F(2.0) / (r * r + i * i + j * j + k * k)).astype(F)                   # rc:262-263
    m00 = F(1) - two_s * (j * j + k * k)                                           # rc:265-277
    m01 = two_s * (i * j - k * r)
    m02 = two_s * (i * k + j * r)
    m12 = two_s * (j * k - i * r)
    m22 = F(1) - two_s * (i * i + j * j)
    # rc:342-385 with convention 'XYZ': i0 = 0, i2 = 2, Tait-Bryan; central angle asin(+M[0,2]) (rc:367-370);
    # first angle from column 2 -> atan2(-M[1,2], M[2,2]); third from row 0 -> atan2(-M[0,1], M[0,0]) (rc:299-329)
    e0 = np.arctan2(-m12, m22)
    e1 = np.arcsin(m02)
    e2 = np.arctan2(-m01, m00)
    return np.stack([e0, e1, e2], -1).astype(F)


def beat_axis_angle_branch(motion_aa, mean_aa, std_aa, mean_pose, std_pose):
    """beat:1056-1062.  ``motion_aa [B,T,C]`` normalised axis-angle gesture (C = 3 * joints).
    Returns (euler_out in degrees [B,T,C] -- what ``result2target_vis`` writes to BVH, beat:1076 --, out_motions normalised)."""
    x = np.asarray(motion_aa, F)
    B, T, C = x.shape
    denorm = (x * np.asarray(std_aa, F)).astype(F) + np.asarray(mean_aa, F)        # beat:1057
    euler = axis_angle_to_euler_xyz(denorm.reshape(B, T, C // 3, 3)).reshape(B, T, C)
    euler = (euler * F(180 / np.pi)).astype(F)                                     # beat:1060
    out = ((euler - np.asarray(mean_pose, F)) / np.asarray(std_pose, F)).astype(F)  # beat:1061
    return euler, out
