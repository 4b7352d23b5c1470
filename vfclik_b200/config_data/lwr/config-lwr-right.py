# Robot configuration: KUKA LWR 4+, right arm, 7 DOF.
#
# The reference loads "<config_dir>config-<robot>-<instance>.py" (scripts/vfclik:80-85)
# from an external config_data repository that is not part of arcoslab/vfclik.  This
# file supplies the attributes the hot path reads with the nominal
# LWR geometry: link offsets 0.31 / 0.40 / 0.39 / 0.078 m, alternating +-pi/2 twists,
# joint limits +-170 / +-120 degrees alternating.  The geometry is this repo's choice
# (ORACLE_CHOICE "chain constants"): any KDL-style segment list works.
from math import pi

from vfclik_b200.kdl import Frame, Joint, Rotation, Segment, Vector

arm_type = "lwr"
arm_instance = "right"
robotarm_portbasename = "/lwr/right"

nJoints = 7

# Mounting of the arm base in the world frame.  The reference's first goal for this arm
# (old/system_start.sh.old:346, position 0.8 / -0.15 / 1.2 m) lies outside the reach of a
# floor-mounted LWR, i.e. the reference robot's arm sits on a raised torso; this offset
# (repo's choice) puts that goal inside the workspace.
mounting = Frame(Rotation.Identity(), Vector(0.25, 0.0, 0.45))

segments = [
    Segment(Joint(Joint.NoJoint), mounting),
    Segment(Joint(Joint.NoJoint), Frame.DH_Craig1989(0.0, 0.0, 0.31, 0.0)),
    Segment(Joint(Joint.RotZ), Frame.DH_Craig1989(0.0, pi / 2, 0.0, 0.0)),
    Segment(Joint(Joint.RotZ), Frame.DH_Craig1989(0.0, -pi / 2, 0.40, 0.0)),
    Segment(Joint(Joint.RotZ), Frame.DH_Craig1989(0.0, -pi / 2, 0.0, 0.0)),
    Segment(Joint(Joint.RotZ), Frame.DH_Craig1989(0.0, pi / 2, 0.39, 0.0)),
    Segment(Joint(Joint.RotZ), Frame.DH_Craig1989(0.0, pi / 2, 0.0, 0.0)),
    Segment(Joint(Joint.RotZ), Frame.DH_Craig1989(0.0, -pi / 2, 0.0, 0.0)),
    Segment(Joint(Joint.RotZ), Frame(Rotation.Identity(), Vector(0.0, 0.0, 0.078))),
]

_deg = pi / 180.0
limits = [[-170 * _deg, 170 * _deg], [-120 * _deg, 120 * _deg]] * 3 + [[-170 * _deg, 170 * _deg]]


def updateJntLimits(q):
    """Position-dependent limits hook (scripts/joint_p_controller:80); constant for the LWR."""
    return limits


# start posture and first goal: old/system_start.sh.old:234 and :346
initial_joint_pos = [0.0, -1.2, 0.7, 1.4, 0.35, -1.4, 0.0]
initial_vf_pose = ["set", "goal",
                   [-1, 0, 0, 0.8, 0, 0, 1, -0.15, 0, 1, 0, 1.2, 0, 0, 0, 1, 0.05]]

# gains / rates
speedScale = 0.2          # Cartesian speed scale, cap 0.41 (scripts/vf:134-137,200)
jpctrl_kp = 1.5           # scripts/joint_p_controller:55-56
max_vel = 1.0             # rad/s, leading-joint clamp (scripts/bridge:69,188-196)
rate = 0.01               # s, bridge period (scripts/bridge:91,629-634)

# velocity-IK / nullspace parameters (explicit here; hidden inside Lafik in the reference)
ik_lambda = 0.1           # damping of J^T (J J^T + lambda^2 I)^-1
ns_lambda = 0.1           # damping of the nullspace projector's pseudo-inverse (0 = pinv)
ns_limit_gain = 1.0       # k of the joint-limit-avoidance gradient qdot0

# bridge plumbing (scripts/bridge:122-132)
torso_joints = []
qin_portname = "/bridge/qin"
qcmded_portname = "/bridge/qcmded"
qcmd_portname = "/bridge/qcmd"
