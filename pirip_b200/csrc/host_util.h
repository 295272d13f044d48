/* host_util.h — error reporting and launch accounting shared by the host translation units */
#ifndef PIRIP_HOST_UTIL_H
#define PIRIP_HOST_UTIL_H

#include <cuda_runtime.h>

#include <string>

namespace pirip {

void set_error(const std::string &msg);
/* records msg, returns code */
int fail(int code, const std::string &msg);
/* true when e == cudaSuccess; otherwise records "<what>: <cuda error string>" */
bool cuda_ok(cudaError_t e, const char *what);
void count_launch(unsigned n = 1);
int current_device();
const char *last_error();

}  // namespace pirip

#endif
