"""CPU ORACLE (test infrastructure, NOT product code) -- tracking diagnostics and distance monitor.

Reference-shaped scalar restatement, one instance at a time:
``TrackingDiagnostics.update`` follows ``scripts/vf:349-428`` (frame list of 5, command buffer of 4, check delay 4,
loop_freq 150, tracking threshold 0.10); ``DistanceMonitor.update`` follows ``scripts/monitor_distance:76-84,148-219``
(orientLength in degrees, thresholds 0.02 m / 1 degree / 0.1 rad, majority over a 20-sample buffer).  Both use
``kdl_diff`` from oracle/refshape.py (PARITY UNPINNED for the PyKDL part, see oracle/batch.py).
"""
from __future__ import annotations

from math import acos, pi, sqrt

import numpy as np

from .refshape import KdlFrame, kdl_diff

STATES = ["on goal", "follow", "not follow"]


class TrackingDiagnostics:
    def __init__(self):
        self.frame_list, self.cmd_buffer = [], []
        self.cmd_buffer_size, self.frame_list_size, self.check_delay = 4, 5, 4

    def update(self, frame: KdlFrame, velPos, velRot):
        """Returns the 8-tuple of /vectorField/track_error, or None while the buffers fill (scripts/vf:349-428)."""
        velPos, velRot = np.asarray(velPos, dtype=np.float64), np.asarray(velRot, dtype=np.float64)
        self.cmd_buffer.append([velPos, velRot])
        if len(self.cmd_buffer) > self.cmd_buffer_size:
            self.cmd_buffer.pop(0)
        self.frame_list.append(frame)
        if len(self.frame_list) <= self.frame_list_size:
            return None
        self.frame_list.pop(0)
        ext_diff = kdl_diff(self.frame_list[len(self.frame_list) - 2], self.frame_list[len(self.frame_list) - 1])
        cmd = self.cmd_buffer[len(self.cmd_buffer) - self.check_delay]
        ext_vel_mag = sqrt(float(ext_diff.vel @ ext_diff.vel))
        cmd_vel_mag = sqrt(float(cmd[0] @ cmd[0]))
        ext_vel = ext_diff.vel / ext_vel_mag if ext_vel_mag > 0 else np.array([1.0, 0, 0])
        cmd_vel = cmd[0] / cmd_vel_mag if cmd_vel_mag > 0 else np.array([1.0, 0, 0])
        vel_dot = min(1.0, max(-1.0, float(cmd_vel @ ext_vel)))
        vel_diff_angle = sqrt(acos(vel_dot) ** 2)
        ext_rot_mag = sqrt(float(ext_diff.rot @ ext_diff.rot))
        cmd_rot_mag = sqrt(float(cmd[1] @ cmd[1]))
        ext_rot = ext_diff.rot / ext_rot_mag if ext_rot_mag > 0 else np.array([1.0, 0, 0])
        cmd_rot = cmd[1] / cmd_rot_mag if cmd_rot_mag > 0 else np.array([1.0, 0, 0])
        rot_dot = min(1.0, max(-1.0, float(cmd_rot @ ext_rot)))
        rot_diff_angle = sqrt(acos(rot_dot) ** 2)
        loop_freq, tracking_th = 150, 0.10
        ext_vel_mag_corr, ext_rot_mag_corr = ext_vel_mag * loop_freq, ext_rot_mag * loop_freq
        cmd_rot_mag_corr, cmd_vel_mag_corr = cmd_rot_mag / 5.0, cmd_vel_mag * 1.0
        ext_int_diff = abs((cmd_vel_mag_corr + cmd_rot_mag_corr) - (ext_vel_mag_corr + ext_rot_mag_corr))
        return (vel_diff_angle, rot_diff_angle, ext_vel_mag_corr, ext_rot_mag_corr, cmd_vel_mag_corr, cmd_rot_mag_corr,
                ext_int_diff, int(ext_int_diff < tracking_th))


def orientLength(final: KdlFrame, current: KdlFrame) -> float:
    """scripts/monitor_distance:76-84: rotation angle between the frames, in degrees."""
    tw = kdl_diff(current, final)
    return 180.0 * sqrt(tw.rot[0] ** 2 + tw.rot[1] ** 2 + tw.rot[2] ** 2) / pi


class DistanceMonitor:
    def __init__(self):
        self.track_error_xyz = self.track_error_rot = 0.0
        self.tracking_buffer = []
        self.tracking_buffer_size = 20

    def update(self, frame: KdlFrame, goal: KdlFrame, track_error):
        """Returns (distanceXYZ, distanceOrient_deg, xyz_majority | None, rot_majority | None)."""
        if track_error is not None:
            self.track_error_xyz, self.track_error_rot = track_error[0], track_error[1]
        d = frame.p - goal.p
        distanceXYZ = sqrt(float(d @ d))
        distanceOrient = orientLength(goal, frame)
        xyz_state = rot_state = "on goal"
        if distanceXYZ > 0.02 and self.track_error_xyz > 0.1:
            xyz_state = "not follow"
        if distanceXYZ > 0.02 and self.track_error_xyz < 0.1:
            xyz_state = "follow"
        if distanceXYZ < 0.02:
            xyz_state = "on goal"
        if distanceOrient > 1.0 and self.track_error_rot > 0.1:
            rot_state = "not follow"
        if distanceOrient > 1.0 and self.track_error_rot < 0.1:
            rot_state = "follow"
        if distanceOrient < 1.0:
            rot_state = "on goal"
        self.tracking_buffer.append([xyz_state, rot_state])
        mx = mr = None
        if len(self.tracking_buffer) > self.tracking_buffer_size:
            self.tracking_buffer.pop(0)
            cols = list(zip(*self.tracking_buffer))
            mx = max(STATES, key=list(cols[0]).count)
            mr = max(STATES, key=list(cols[1]).count)
        return distanceXYZ, distanceOrient, mx, mr
