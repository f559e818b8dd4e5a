/* Hand-written stand-in for the autoconf-generated config.h the reference
 * expects (src/config.h.in is absent from the tree).  TEST INFRASTRUCTURE:
 * used only to compile the unmodified reference from /root/reference/src
 * into oracle/_ref/.  See interface.c:375,384 and kslib.h:111-135. */
#ifndef TDG_ORACLE_CONFIG_H
#define TDG_ORACLE_CONFIG_H
#define PACKAGE_NAME "Tagdust"
#define PACKAGE_VERSION "2.33"
#define PACKAGE_BUGREPORT "timolassmann@gmail.com"
#define kslDEBUGLEVEL 0
#endif
