// C-ABI trampoline around the reference's `cufd` (Src/libCUFD.h:6-10), whose
// last argument is a std::string by value and therefore not callable from
// ctypes.  Ours; compiled together with the reference sources by oracle/Makefile
// into oracle/_ref/libcufd_ref.so.  Test infrastructure only.
#include <string>
extern "C" void cufd(float *misfit, float *grad_Lambda, float *grad_Mu, float *grad_Den,
                     float *grad_stf, const float *Lambda, const float *Mu, const float *Den,
                     const float *stf, int calc_id, const int gpu_id, const int group_size,
                     const int *shot_ids, const std::string para_fname);

extern "C" void ref_cufd(float *misfit, float *grad_Lambda, float *grad_Mu, float *grad_Den,
                         float *grad_stf, const float *Lambda, const float *Mu, const float *Den,
                         const float *stf, int calc_id, int gpu_id, int group_size,
                         const int *shot_ids, const char *para_fname)
{
    cufd(misfit, grad_Lambda, grad_Mu, grad_Den, grad_stf, Lambda, Mu, Den, stf, calc_id, gpu_id,
         group_size, shot_ids, std::string(para_fname));
}
