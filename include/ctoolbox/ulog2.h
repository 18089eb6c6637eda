/*
 * ctoolbox/ulog2.h -- floor(log2(x)) for a non-zero 32 bit value; the single
 * arithmetic helper the reference takes from ctoolbox (used in the lazy-match
 * accept rule, reference src/deflator.c:2876-2877).
 */
#ifndef JDB200_CTOOLBOX_ULOG2_H
#define JDB200_CTOOLBOX_ULOG2_H

#include "ctoolbox.h"

static inline __attribute__((unused)) uint32
ctb_u32log2(uint32 v)
{
	return v ? (uint32) (31 - __builtin_clz(v)) : 0u;
}

#endif
