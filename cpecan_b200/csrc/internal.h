/* internal.h -- shared between the C host files and the CUDA engine of libcpecan_b200. */
#ifndef CPB_INTERNAL_H_
#define CPB_INTERNAL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void cpb_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
