// placeholder: replaced by the CPU HNSW baseline (timing only) in a later commit
extern "C" int orc_hnsw_placeholder(void) { return 0; }
