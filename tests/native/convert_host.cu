// Host build of the converter recursions (csrc/convert_row.cuh) for tests/test_convert_host.py.
#include "convert_row.cuh"

extern "C" {
void convert_rows_host_f32(float* a, long rows, int D, int op, double g) {
  for (long r = 0; r < rows; ++r) dsb200::convert_row<float>(a + r * D, D, op, static_cast<float>(g));
}
void convert_rows_host_f64(double* a, long rows, int D, int op, double g) {
  for (long r = 0; r < rows; ++r) dsb200::convert_row<double>(a + r * D, D, op, g);
}
}
