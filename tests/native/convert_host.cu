// Host build of the converter recursions (csrc/convert_row.cuh) for tests/test_convert_host.py.
#include "convert_row.cuh"

template <typename T>
static void rows_fixed(T* a, long rows, int D, int op, T g) {
  for (long r = 0; r < rows; ++r) {
    if (D == 13) dsb200::convert_row_fixed<T, 13>(a + r * D, op, g);
    else if (D == 25) dsb200::convert_row_fixed<T, 25>(a + r * D, op, g);
    else if (D == 3) dsb200::convert_row_fixed<T, 3>(a + r * D, op, g);
    else if (D == 2) dsb200::convert_row_fixed<T, 2>(a + r * D, op, g);
    else if (D == 8) dsb200::convert_row_fixed<T, 8>(a + r * D, op, g);
    else dsb200::convert_row<T>(a + r * D, D, op, g);
  }
}

extern "C" {
void convert_rows_fixed_host_f32(float* a, long rows, int D, int op, double g) { rows_fixed<float>(a, rows, D, op, static_cast<float>(g)); }
void convert_rows_fixed_host_f64(double* a, long rows, int D, int op, double g) { rows_fixed<double>(a, rows, D, op, g); }
void convert_rows_host_f32(float* a, long rows, int D, int op, double g) {
  for (long r = 0; r < rows; ++r) dsb200::convert_row<float>(a + r * D, D, op, static_cast<float>(g));
}
void convert_rows_host_f64(double* a, long rows, int D, int op, double g) {
  for (long r = 0; r < rows; ++r) dsb200::convert_row<double>(a + r * D, D, op, g);
}
}
