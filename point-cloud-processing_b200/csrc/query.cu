// kNN rows, radius search, density filter, scans: part 0 of query_body.inc
#define PCPX_QUERY_PART 0
#include "query_body.inc"
