/* /root/reference/mobilevit/main.cpp:4 includes "common.h" (from upstream ggml's examples/) and uses nothing
 * from it; this empty header keeps an unmodified main.cpp compiling against libggml_b200. */
#ifndef GGML_B200_COMMON_H
#define GGML_B200_COMMON_H
#endif
