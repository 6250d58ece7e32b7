// aud_internal.h -- declarations shared by the host-side translation units.
#ifndef AUD_INTERNAL_H_
#define AUD_INTERNAL_H_

#include <cstdint>

namespace aud {

// Record `msg` as the calling thread's last error and return `code`.
int32_t fail(int32_t code, const char *msg);
int32_t failf(int32_t code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

}  // namespace aud

#endif  // AUD_INTERNAL_H_
