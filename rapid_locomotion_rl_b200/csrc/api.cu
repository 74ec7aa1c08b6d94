// C-ABI plumbing: error text, version, launch checks.
#include <stdarg.h>
#include <string.h>

#include "rl_common.cuh"

namespace rl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t err = cudaPeekAtLastError();
  if (err != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(err));
    (void)cudaGetLastError();
    return RL_ERR_CUDA;
  }
  return RL_OK;
}

}  // namespace rl

extern "C" const char* rl_last_error(void) { return rl::g_err; }
extern "C" const char* rl_version(void) { return "rl_b200 0.1.0 (sm_100a)"; }

// struct sizes, so a binding generated from the header can verify its layout at load time
extern "C" int64_t rl_sizeof(const char* name) {
  if (!strcmp(name, "RlEnvCfg")) return sizeof(RlEnvCfg);
  if (!strcmp(name, "RlEnvBuffers")) return sizeof(RlEnvBuffers);
  if (!strcmp(name, "RlResetCfg")) return sizeof(RlResetCfg);
  if (!strcmp(name, "RlResetBuffers")) return sizeof(RlResetBuffers);
  if (!strcmp(name, "RlGacCfg")) return sizeof(RlGacCfg);
  if (!strcmp(name, "RlGacBuffers")) return sizeof(RlGacBuffers);
  if (!strcmp(name, "RlPeerComm")) return sizeof(RlPeerComm);
  if (!strcmp(name, "RlStorageAdd")) return sizeof(RlStorageAdd);
  if (!strcmp(name, "RlWgradProblem")) return sizeof(RlWgradProblem);
  if (!strcmp(name, "RlChainTensor")) return sizeof(RlChainTensor);
  if (!strcmp(name, "RlChainLoadOp")) return sizeof(RlChainLoadOp);
  if (!strcmp(name, "RlChainMmaOp")) return sizeof(RlChainMmaOp);
  if (!strcmp(name, "RlChainEpiOp")) return sizeof(RlChainEpiOp);
  if (!strcmp(name, "RlChainDesc")) return sizeof(RlChainDesc);
  if (!strcmp(name, "RlChainPpoLoss")) return sizeof(RlChainPpoLoss);
  if (!strcmp(name, "RlRolloutBoundary")) return sizeof(RlRolloutBoundary);
  if (!strcmp(name, "RlRolloutAct")) return sizeof(RlRolloutAct);
  return -1;
}
