// cuipm_internal.h -- declarations shared by the host and device translation units of libcuipm.
#ifndef CUIPM_INTERNAL_H_
#define CUIPM_INTERNAL_H_

#include <string>

#include "cuipm.h"

namespace cuipm {
void set_error(const std::string &msg);
// CUIPM_OK if the option values are within what the device path implements, else CUIPM_ERR_INVALID (+ message)
int opts_check(const cuipm_opts *o);
}  // namespace cuipm

#endif
