// oracle/ref_tfops_shim.cu -- TEST INFRASTRUCTURE ONLY.
// extern "C" entry points around the reference's OWN launchers, which are compiled from the
// sources where they lie under /root/reference (never copied into this repo):
//   pointnet_plusplus/utils/tf_ops/sampling/tf_sampling_g.cu:203-208
//   pointnet_plusplus/utils/tf_ops/grouping/tf_grouping_g.cu:125-135
// The launchers use the legacy default stream and never check errors; the shims synchronise and
// return cudaGetLastError() so tests can use them as a bit-exact GPU-side oracle and bench.py can
// time them as the incumbent kernels.  All pointers are DEVICE pointers.
#include <cuda_runtime.h>

void farthestpointsamplingLauncher(int b, int n, int m, const float *inp, float *temp, int *out);
void gatherpointLauncher(int b, int n, int m, const float *inp, const int *idx, float *out);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                            int *idx, int *pts_cnt);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out);

extern "C" {
// temp must hold 32*n floats (tf_sampling.cpp:115)
int ref_fps(int b, int n, int m, const float *inp, float *temp, int *out, int sync)
{
    farthestpointsamplingLauncher(b, n, m, inp, temp, out);
    if (sync) cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}
int ref_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, int sync)
{
    gatherpointLauncher(b, n, m, inp, idx, out);
    if (sync) cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}
int ref_ball_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                   int *pts_cnt, int sync)
{
    queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
    if (sync) cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}
int ref_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                    int sync)
{
    groupPointLauncher(b, n, c, m, nsample, points, idx, out);
    if (sync) cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}
}
