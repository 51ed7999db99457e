/* ORACLE (test infrastructure, never shipped): regex matcher restating the
 * `regex 1.12.2` crate's `Regex::is_match` on ASCII haystacks.
 * The crate is not vendored in the reference (Cargo.lock:1694-1695); this follows
 * its published syntax (regex-syntax 0.8) and semantics: leftmost search,
 * no look-around / back-references, `$` = end of haystack, `.` excludes \n. */
#ifndef ORACLE_RX_H
#define ORACLE_RX_H
#include <stddef.h>
#include <stdint.h>

enum { RX_OK = 0, RX_INVALID = 1, RX_UNSUPPORTED = 2, RX_TOO_BIG = 3 };

typedef struct rx_prog rx_prog;

/* Compile; *status gets RX_*; returns NULL unless RX_OK.  `err` (>=128 bytes) receives a message. */
rx_prog* rx_compile(const char* pat, size_t len, int* status, char* err);
int rx_is_match(const rx_prog* p, const uint8_t* s, size_t n);
void rx_free(rx_prog* p);

#endif
