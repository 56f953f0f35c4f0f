// Error convention of the drop-in headers: Jet::Exception carrying
// "[file][Line:n][Method:f]: Error in Jet: msg" like the reference
// (/root/reference/include/jet/Abort.hpp:51-97); failures of the CUDA library surface through
// the same exception with the text of jb_last_error().
#pragma once

#include <exception>
#include <sstream>
#include <string>

namespace Jet {

class Exception : public std::exception {
  public:
    explicit Exception(const std::string &what_arg) : message_(what_arg) {}
    explicit Exception(const char *what_arg) : message_(what_arg) {}
    const char *what() const noexcept override { return message_.c_str(); }

  private:
    std::string message_;
};

[[noreturn]] inline void Abort(const std::string &message, const char *file, int line,
                               const char *function)
{
    std::ostringstream os;
    os << "[" << file << "][Line:" << line << "][Method:" << function
       << "]: Error in Jet: " << message;
    throw Exception(os.str());
}

} // namespace Jet

#define JET_ABORT(message) ::Jet::Abort((message), __FILE__, __LINE__, __func__)
#define JET_ABORT_IF(cond, message)                                                               \
    do {                                                                                          \
        if (cond) {                                                                               \
            JET_ABORT(message);                                                                   \
        }                                                                                         \
    } while (0)
#define JET_ABORT_IF_NOT(cond, message) JET_ABORT_IF(!(cond), message)
#define JET_ASSERT(cond) JET_ABORT_IF_NOT(cond, "Assertion failed: " #cond)
// Status of a jb_* call -> exception (the analogue of JET_CUDA_IS_SUCCESS,
// /root/reference/include/jet/CudaTensorHelpers.hpp:18-59)
#define JET_JB_CHECK(expr)                                                                        \
    do {                                                                                          \
        if ((expr) != 0) {                                                                        \
            JET_ABORT(jb_last_error());                                                           \
        }                                                                                         \
    } while (0)
