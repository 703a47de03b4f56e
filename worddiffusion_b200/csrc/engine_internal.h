// Internals shared between the translation units of libwd_b200.so (not part of the C ABI).
#pragma once
// records `msg` as the calling thread's wd_last_error() and returns `code`
int wd_set_error(int code, const char* msg);
