/* the reference links GNU libiconv; glibc's iconv has the same semantics */
#include <iconv.h>
#include <stddef.h>
void* libiconv_open(const char* to, const char* from) { return (void*)iconv_open(to, from); }
size_t libiconv(void* cd, char** in, size_t* inleft, char** out, size_t* outleft)
{ return iconv((iconv_t)cd, in, inleft, out, outleft); }
int libiconv_close(void* cd) { return iconv_close((iconv_t)cd); }
