/* ORACLE SHIM -- test infrastructure.  Stand-in for xobjects/headers/atomicadd.h
 * (included by xtrack/headers/track.h:10); nothing on the single-particle hot
 * path calls atomicAdd when in-kernel record logging is off. */
#ifndef XTB_ORACLE_XO_ATOMICADD_H
#define XTB_ORACLE_XO_ATOMICADD_H
#endif
