// soap.cu -- placeholder until the SOAP kernels land (next commit).
#include "common.cuh"
using namespace dsg;
extern "C" int dsg_soap_n_features(int n_species, int n_max, int l_max) {
    if (n_species < 1 || n_max < 1 || l_max < 0) return DSG_ERR_INVALID_ARGUMENT;
    const long long sn = (long long)n_species * n_max;
    return (int)(sn * (sn + 1) / 2 * (l_max + 1));
}
extern "C" int dsg_soap_workspace_bytes(int, int, int, int, int, int, const double*, double,
                                        size_t* bytes_out) {
    if (bytes_out) *bytes_out = 0;
    return set_error(DSG_ERR_UNSUPPORTED, "SOAP kernels not built yet");
}
extern "C" int dsg_soap(const void*, int, const int32_t*, const int32_t*, int, int, int, int,
                        const double*, double, int, int, const double*, const double*, void*,
                        size_t, double*, int64_t, int64_t, void*) {
    return set_error(DSG_ERR_UNSUPPORTED, "SOAP kernels not built yet");
}
