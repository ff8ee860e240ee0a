// internal to libbin3c_io.so: the thread-local message behind b3c_io_last_error()
#pragma once
namespace b3cio {
void set_err(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
}
