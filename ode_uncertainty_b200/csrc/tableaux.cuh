// Butcher tableaux of the explicit embedded Runge-Kutta solvers, as compile-time constants.
//
// Values restate the reference's tableau registry (they are mathematical constants):
//   RKF45     src/solvers/rkf45.py:10-34        (S=6)
//   Dopri65   src/solvers/dopri65.py:10-72      (S=8)
//   BS32      src/solvers/bs32.py:10-32         (S=4)
//   HeunEuler src/solvers/heun_euler.py:10-30   (S=2; row b[1] = [0.5, 0] sums to 0.5 in the
//                                                reference and is kept verbatim)
// Convention (src/solvers/rksolver.py:63-64,146-151): row b[1] PROPAGATES the state, row b[0]
// is used only for the embedded error estimate eps = |x_next(b[0]) - x_next(b[1])|.
//
// The kernels unroll every stage loop, so a(i,j)/b(r,j)/c(i) with literal indices fold to
// constant-bank operands and structural zeros disappear from the instruction stream.
#pragma once

namespace odeu {

struct TabRKF45 {
  static constexpr int S = 6;
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double A[6][6] = {
        {0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
        {1.0 / 4, 0.0, 0.0, 0.0, 0.0, 0.0},
        {3.0 / 32, 9.0 / 32, 0.0, 0.0, 0.0, 0.0},
        {1932.0 / 2197, -7200.0 / 2197, 7296.0 / 2197, 0.0, 0.0, 0.0},
        {439.0 / 216, -8.0, 3680.0 / 513, -845.0 / 4104, 0.0, 0.0},
        {-8.0 / 27, 2.0, -3544.0 / 2565, 1859.0 / 4104, -11.0 / 40, 0.0}};
    return A[i][j];
  }
  __host__ __device__ static constexpr double b(int r, int j) {
    constexpr double Bm[2][6] = {
        {16.0 / 135, 0.0, 6656.0 / 12825, 28561.0 / 56430, -9.0 / 50, 2.0 / 55},
        {25.0 / 216, 0.0, 1408.0 / 2565, 2197.0 / 4104, -1.0 / 5, 0.0}};
    return Bm[r][j];
  }
  __host__ __device__ static constexpr double c(int i) {
    constexpr double C[6] = {0.0, 1.0 / 4, 3.0 / 8, 12.0 / 13, 1.0, 1.0 / 2};
    return C[i];
  }
};

struct TabDopri65 {
  static constexpr int S = 8;
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double A[8][8] = {
        {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
        {1.0 / 10, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
        {-2.0 / 81, 20.0 / 81, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
        {615.0 / 1372, -270.0 / 343, 1053.0 / 1372, 0.0, 0.0, 0.0, 0.0, 0.0},
        {3243.0 / 5500, -54.0 / 55, 50949.0 / 71500, 4998.0 / 17875, 0.0, 0.0, 0.0, 0.0},
        {-26492.0 / 37125, 72.0 / 55, 2808.0 / 23375, -24206.0 / 37125, 338.0 / 459, 0.0, 0.0, 0.0},
        {5561.0 / 2376, -35.0 / 11, -24117.0 / 31603, 899983.0 / 200772, -5225.0 / 1836,
         3925.0 / 4056, 0.0, 0.0},
        {465467.0 / 266112, -2945.0 / 1232, -5610201.0 / 14158144, 10513573.0 / 3212352,
         -424325.0 / 205632, 376225.0 / 454272, 0.0, 0.0}};
    return A[i][j];
  }
  __host__ __device__ static constexpr double b(int r, int j) {
    constexpr double Bm[2][8] = {
        {821.0 / 10800, 0.0, 19683.0 / 71825, 175273.0 / 912600, 395.0 / 3672, 785.0 / 2704,
         3.0 / 50, 0.0},
        {61.0 / 864, 0.0, 98415.0 / 321776, 16807.0 / 146016, 1375.0 / 7344, 1375.0 / 5408,
         -37.0 / 1120, 1.0 / 10}};
    return Bm[r][j];
  }
  __host__ __device__ static constexpr double c(int i) {
    constexpr double C[8] = {0.0, 1.0 / 10, 2.0 / 9, 3.0 / 7, 3.0 / 5, 4.0 / 5, 1.0, 1.0};
    return C[i];
  }
};

struct TabBS32 {
  static constexpr int S = 4;
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double A[4][4] = {{0.0, 0.0, 0.0, 0.0},
                                {1.0 / 2, 0.0, 0.0, 0.0},
                                {0.0, 3.0 / 4, 0.0, 0.0},
                                {2.0 / 9, 1.0 / 3, 4.0 / 9, 0.0}};
    return A[i][j];
  }
  __host__ __device__ static constexpr double b(int r, int j) {
    constexpr double Bm[2][4] = {{7.0 / 24, 1.0 / 4, 1.0 / 3, 1.0 / 8},
                                 {2.0 / 9, 1.0 / 3, 4.0 / 9, 0.0}};
    return Bm[r][j];
  }
  __host__ __device__ static constexpr double c(int i) {
    constexpr double C[4] = {0.0, 1.0 / 2, 3.0 / 4, 1.0};
    return C[i];
  }
};

struct TabHeunEuler {
  static constexpr int S = 2;
  __host__ __device__ static constexpr double a(int i, int j) {
    constexpr double A[2][2] = {{0.0, 0.0}, {1.0, 0.0}};
    return A[i][j];
  }
  __host__ __device__ static constexpr double b(int r, int j) {
    constexpr double Bm[2][2] = {{0.5, 0.5}, {0.5, 0.0}};
    return Bm[r][j];
  }
  __host__ __device__ static constexpr double c(int i) {
    constexpr double C[2] = {0.0, 1.0};
    return C[i];
  }
};

}  // namespace odeu
