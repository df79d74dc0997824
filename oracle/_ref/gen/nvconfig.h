#ifndef NV_CONFIG
#define NV_CONFIG
#define HAVE_UNISTD_H
#define HAVE_STDARG_H
#define HAVE_SIGNAL_H
#define HAVE_EXECINFO_H
#define HAVE_MALLOC_H
#define NV_HAVE_STBIMAGE
#endif
