// placeholder, replaced by the tcgen05 engine
#include "dn_internal.h"
bool tc_supported_device() { return false; }
int tc_rows_chain(const DnRowsSrc&, const DnLayer*, int, int64_t, int, void*, int64_t, cudaStream_t) { return DN_ERR_UNSUPPORTED; }
int tc_rows_chain_supported(const DnRowsSrc&, const DnLayer*, int) { return DN_ERR_UNSUPPORTED; }
int tc_to_basis_partial(const float*, const float*, const float*, int64_t, int, int, float*, int*, int, cudaStream_t) { return DN_ERR_UNSUPPORTED; }
int tc_to_basis_supported(int, int) { return DN_ERR_UNSUPPORTED; }
int64_t tc_chain_ws_bytes(const DnLayer*, int) { return 0; }
