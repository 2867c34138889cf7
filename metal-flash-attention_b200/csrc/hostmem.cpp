// Host-memory placement for the host-buffer path (mfa_attention_run_host): page-locked buffers on the NUMA node the
// GPU hangs off.  On a two-socket HGX box GPUs 0-3 and 4-7 sit on different sockets; a pinned buffer that the first-touch
// policy happened to place on the other socket makes every H2D / D2H copy cross the inter-socket link, and with all
// eight ranks copying at once that link -- not PCIe -- bounds the end-to-end rate (round-1 scaling: 0.62 at 8 GPUs).
// Everything here is Linux sysfs + sched_setaffinity + cudaHostAlloc: no libnuma dependency.
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "internal.h"

namespace {

bool read_line(const std::string &path, std::string &out) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return false;
  char buf[4096];
  const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
  fclose(f);
  if (!ok) return false;
  out = buf;
  while (!out.empty() && (out.back() == '\n' || out.back() == ' ')) out.pop_back();
  return true;
}

// "0-31,64-95" -> cpu_set_t
bool parse_cpulist(const std::string &list, cpu_set_t &set) {
  CPU_ZERO(&set);
  const char *p = list.c_str();
  bool any = false;
  while (*p) {
    char *end = nullptr;
    long lo = strtol(p, &end, 10);
    if (end == p) return false;
    long hi = lo;
    p = end;
    if (*p == '-') {
      hi = strtol(p + 1, &end, 10);
      if (end == p + 1) return false;
      p = end;
    }
    for (long c = lo; c <= hi && c < CPU_SETSIZE; ++c) {
      CPU_SET(static_cast<int>(c), &set);
      any = true;
    }
    if (*p == ',') ++p;
  }
  return any;
}

// NUMA node of the GPU (from its PCI function's sysfs entry) and the CPUs of that node the process may run on
int device_numa_cpus(int device, int *node_out, cpu_set_t *cpus_out) {
  char busid[32] = {};
  cudaError_t e = cudaDeviceGetPCIBusId(busid, sizeof(busid), device);
  if (e != cudaSuccess)
    return mfa::fail(MFA_ERROR_NO_DEVICE, std::string("cudaDeviceGetPCIBusId: ") + cudaGetErrorString(e));
  for (char *c = busid; *c; ++c) *c = static_cast<char>(tolower(*c));
  std::string text;
  int node = -1;
  if (read_line(std::string("/sys/bus/pci/devices/") + busid + "/numa_node", text)) node = atoi(text.c_str());
  *node_out = node;
  if (node < 0) return MFA_SUCCESS;  // single-node machine or the platform does not say: nothing to bind to
  cpu_set_t node_cpus, allowed;
  if (!read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", text) ||
      !parse_cpulist(text, node_cpus)) {
    *node_out = -1;
    return MFA_SUCCESS;
  }
  // respect the cgroup / taskset limits the process already runs under
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    cpu_set_t both;
    CPU_AND(&both, &node_cpus, &allowed);
    if (CPU_COUNT(&both) > 0) node_cpus = both;
    else *node_out = -1;  // none of the node's CPUs is available to this process: leave the affinity alone
  }
  *cpus_out = node_cpus;
  return MFA_SUCCESS;
}

}  // namespace

extern "C" {

int mfa_host_bind_thread_to_device(int device, int *numa_node) {
  int node = -1;
  cpu_set_t cpus;
  CPU_ZERO(&cpus);
  int status = device_numa_cpus(device, &node, &cpus);
  if (status != MFA_SUCCESS) return status;
  if (numa_node) *numa_node = node;
  if (node < 0) return MFA_SUCCESS;
  if (sched_setaffinity(0, sizeof(cpus), &cpus) != 0)
    return mfa::fail(MFA_ERROR_INVALID_ARGUMENT, std::string("sched_setaffinity: ") + strerror(errno));
  return MFA_SUCCESS;
}

static int host_alloc(size_t bytes, int device, unsigned extra_flags, void **out) {
  if (!out || bytes == 0) return mfa::fail(MFA_ERROR_INVALID_ARGUMENT, "NULL argument or zero size.");
  int node = -1;
  cpu_set_t cpus, previous;
  CPU_ZERO(&cpus);
  int status = device_numa_cpus(device, &node, &cpus);
  if (status != MFA_SUCCESS) return status;
  // allocate (and thereby first-touch: cudaHostAlloc populates and pins the pages) while running on the GPU's node
  const bool rebound = node >= 0 && sched_getaffinity(0, sizeof(previous), &previous) == 0 &&
                       sched_setaffinity(0, sizeof(cpus), &cpus) == 0;
  int prior_device = -1;
  cudaGetDevice(&prior_device);
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaHostAlloc(out, bytes, cudaHostAllocPortable | extra_flags);
  if (prior_device >= 0 && prior_device != device) cudaSetDevice(prior_device);
  if (rebound) sched_setaffinity(0, sizeof(previous), &previous);
  if (e != cudaSuccess) return mfa::fail(MFA_ERROR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return MFA_SUCCESS;
}

int mfa_host_alloc(size_t bytes, int device, void **out) { return host_alloc(bytes, device, 0u, out); }
int mfa_host_alloc_upload(size_t bytes, int device, void **out) {
  return host_alloc(bytes, device, cudaHostAllocWriteCombined, out);
}

int mfa_host_free(void *ptr) {
  if (!ptr) return MFA_SUCCESS;
  cudaError_t e = cudaFreeHost(ptr);
  if (e != cudaSuccess) return mfa::fail(MFA_ERROR_CUDA, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
  return MFA_SUCCESS;
}

}  // extern "C"
