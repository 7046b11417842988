// fused kNN -> PCA normal kernels: part 1 of query_body.inc
#define PCPX_QUERY_PART 1
#include "query_body.inc"
