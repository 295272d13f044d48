/*
 * Shim for building pirip's own tx/frame_repeater.c (TEST INFRASTRUCTURE, see oracle/Makefile target _ref).
 * The reference source includes codec2's freedv_api_internal.h only for the two rx_status flags it tests
 * ([R tx/frame_repeater.c:29,71,80,88]); codec2 is not vendored in /root/reference, so this header supplies
 * those constants (codec2 freedv_api.h values, the same ones include/pirip_ldpc.h exports) and the libc
 * headers the reference source relies on its includes to pull in.
 */
#ifndef PIRIP_REF_SHIM_FREEDV_API_INTERNAL_H
#define PIRIP_REF_SHIM_FREEDV_API_INTERNAL_H
#include <stdint.h>
#include <string.h>
#define FREEDV_RX_TRIAL_SYNC 0x1
#define FREEDV_RX_SYNC 0x2
#define FREEDV_RX_BITS 0x4
#define FREEDV_RX_BIT_ERRORS 0x8
#endif
