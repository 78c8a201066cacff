"""The EuRoC MAV stereo calibration hard-coded by the reference's driver
(examples/euroc/euroc_example.cpp:96-119): raw intrinsics and distortion of cam0 / cam1, the
rectifying rotations, the rectified projection matrices, the rectified pinhole + baseline handed to
lvt_parameters, and the body-from-camera transform applied to the output trajectory."""
import numpy as np

IMG_SIZE = (752, 480)  # (width, height)

K_L = np.array([[458.654, 0.0, 367.215], [0.0, 457.296, 248.375], [0.0, 0.0, 1.0]])
K_R = np.array([[457.587, 0.0, 379.999], [0.0, 456.134, 255.238], [0.0, 0.0, 1.0]])
P_L = np.array([[435.2046959714599, 0, 367.4517211914062, 0], [0, 435.2046959714599, 252.2008514404297, 0], [0, 0, 1, 0]])
P_R = np.array([[435.2046959714599, 0, 367.4517211914062, -47.90639384423901], [0, 435.2046959714599, 252.2008514404297, 0],
                [0, 0, 1, 0]])
R_L = np.array([[0.999966347530033, -0.001422739138722922, 0.008079580483432283],
                [0.001365741834644127, 0.9999741760894847, 0.007055629199258132],
                [-0.008089410156878961, -0.007044357138835809, 0.9999424675829176]])
R_R = np.array([[0.9999633526194376, -0.003625811871560086, 0.007755443660172947],
                [0.003680398547259526, 0.9999684752771629, -0.007035845251224894],
                [-0.007729688520722713, 0.007064130529506649, 0.999945173484644]])
D_L = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
D_R = np.array([-0.28368365, 0.07451284, -0.00010473, -3.555907e-05, 0.0])

# lvt_parameters after rectification (euroc_example.cpp:109-113)
FX = FY = 435.2046959714599
CX, CY = 367.4517211914062, 252.2008514404297
BASELINE = 0.110077842

# body <- camera (euroc_example.cpp:115-119)
T_BS = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
                 [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
                 [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
                 [0.0, 0.0, 0.0, 1.0]])


def rectify_args(scale=1.0):
    """((K, D, R, P) left, (K, D, R, P) right); scale < 1 shrinks the pinhole for small test images."""
    s = np.diag([scale, scale, 1.0])
    return ((s @ K_L, D_L, R_L, (s @ P_L[:, :3])), (s @ K_R, D_R, R_R, (s @ P_R[:, :3])))
