// A/B switches of the kernels.  The release library has ONE code path per operator: every switch returns its measured-best default and
// no environment variable is read.  A tuning build (HV_BUILD_TUNING=1 python -m humanvid_b200.build -> -DHV_TUNING) lets the scripts under
// scripts/ flip them through HV_* environment variables (README: "Tuning build").
#pragma once
#include <stdlib.h>

namespace hv {
inline long long tune_env(const char* name, long long dflt) {
#ifdef HV_TUNING
  const char* v = getenv(name);
  return v ? atoll(v) : dflt;
#else
  (void)name;
  return dflt;
#endif
}
}  // namespace hv
