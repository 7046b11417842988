// mean kNN distance kernels: part 2 of query_body.inc
#define PCPX_QUERY_PART 2
#include "query_body.inc"
