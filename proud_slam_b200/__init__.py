"""proud_slam_b200 -- B200-native render path for Proud-SLAM (mapping/tracking).

Only what the hot path needs lives here: ``csrc/`` (sm_100a kernels + the C ABI of
``include/proud_slam_b200.h``), the ctypes binding, and the host-side mirror of the
reference's Python interface for this path (``grid``, ``variations.voxel_helpers``,
``variations.render_helpers``, ``variations.nrgbd``, ``criterion``, ``svo``).
"""
__version__ = "0.1.0"
