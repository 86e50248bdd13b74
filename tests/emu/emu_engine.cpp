// TEST INFRASTRUCTURE ONLY -- the WHOLE ENGINE on the host: diffsheg_b200/csrc/engine.cu (handle, packed-weight resolution,
// workspace carving, Runner::prepare_window / Runner::denoise and, through its includes, every kernel they launch) compiled with
// g++ -DDSHEG_EMU -DDSHEG_EMU_RUNTIME.  emu_runtime.h stands in for the CUDA runtime API ("device" pointers are host pointers),
// emu_cuda.h / emu_prims.h / emu_tc_prims.h run each launch on the thread-level emulator.  The exported symbols are the C ABI of
// include/diffsheg_b200.h itself (dsheg_create ... dsheg_denoise), so tests/test_emu_engine.py drives the emulated engine with the
// packer's real output through the same ctypes signatures as the product library.
#include "../../diffsheg_b200/csrc/engine.cu"

extern "C" void emu_engine_set_sms(int n) { emu_rt::num_sms() = n; }
extern "C" long long emu_engine_launches() { return (long long)emu::rt().grids_run; }   // every kernel launch the emulator executed
extern "C" const char* emu_engine_last_launch_error() { return emu_rt::last_launch_error().c_str(); }
extern "C" long long emu_engine_graph_launches() { return emu_rt::graph_launches(); }
