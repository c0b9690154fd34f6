"""articulated_pose_b200 -- B200 (sm_100a) implementation of the ANCSH per-point-cloud hot path.

Host-side mirrors of the reference's interfaces for this path, over the C ABI in include/ancsh_b200.h:

  tf_ops      farthest_point_sample / gather_point / query_ball_point / group_point / three_nn /
              three_interpolate           (pointnet_plusplus/utils/tf_ops/*/tf_*.py wrappers)
  network     AncshNet.forward(P) -> pred dict     (sess.run(pred_dict), lib/network.py:292)
  weights     TF-variable-name keyed weight import  (tf.train.Saver restore, main.py:85-97)

PyTorch is used only for device memory and streams.  There is no CPU fallback: importing the
submodules raises if libancsh_b200.so has not been built.
"""
__all__ = ["tf_ops", "network", "weights", "synthetic"]
