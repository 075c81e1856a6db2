// ABI version of include/danbo_b200.h
extern "C" int danbo_version(void) { return 5; }
