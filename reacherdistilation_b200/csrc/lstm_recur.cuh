// Persistent recurrence kernels of the LSTM student (lstm_recur.cu): the T = 10 LSTMCell(200) steps of
// /root/reference src/distilation/student_nn.py:23,41 forward, and their back-propagation through time, each as ONE launch.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rb {

// private scratch of the two kernels (floats), for B window rows
size_t lstm_recur_ws_floats(int64_t B);

struct LstmRecurArgs {
    const float* W_l;        // [243][800]  rows = [x (43) | m_prev (200)], columns = gates i | j | f | o (200 each)
    const float* b_l;        // [800]
    int64_t B;               // window rows
    float* xh;               // [T*B][256]  in: x columns 0..42 of every step and m_0 in columns 43..242 of step 0; out: m_{t-1} of steps 1..T-1
    float* hh; int hh_ld;    // [T*B][hh_ld >= 200]  out: m_t (input of the per-step heads); columns >= 200 are left alone
    const float* c0;         // [B][200]    initial cell state
    float* c_last;           // [B][200]    out: c_T
    const float* dh;         // [T*B][200]  in (backward): dL/dm_t from the heads
    float* dz;               // [T*B][800]  out (backward): dL/d(pre-activation gates)
    float* dxh;              // [T*B][256]  out (backward): columns 11..42 only (gradient of the prev_pdflat embedding)
    float* scratch;          // lstm_recur_ws_floats(B) floats
};
int lstm_recur_build_images(const LstmRecurArgs& a, cudaStream_t st);     // once per parameter update, before forward / backward
int lstm_recur_forward(const LstmRecurArgs& a, cudaStream_t st);
int lstm_recur_backward(const LstmRecurArgs& a, cudaStream_t st);

}  // namespace rb
