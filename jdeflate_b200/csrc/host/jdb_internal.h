/* jdb_internal.h -- hooks between the host modules (not exported). */
#ifndef JDB_INTERNAL_H
#define JDB_INTERNAL_H

#include <jdeflate/deflator.h>
#include <jdeflate/inflator.h>

/* ask a deflator to checksum the uncompressed bytes it consumes (JDB_CK_* mask,
 * computed on the device-resident batch, no second transfer) */
void jdb_deflator_set_checks(TDeflator* d, int which);
/* running crc register (not finalised) and adler value */
int  jdb_deflator_get_checks(TDeflator* d, uint32* crc, uint32* adler);

/* same for the bytes an inflator produces */
void jdb_inflator_set_checks(TInflator* s, int which);
void jdb_inflator_set_readahead(TInflator* s, size_t bytes);
int  jdb_inflator_get_checks(TInflator* s, uint32* crc, uint32* adler);
/* compressed bytes queued from earlier source windows that lie beyond the end of the stream */
void   jdb_inflator_drop_window(TInflator* s);
size_t jdb_inflator_leftover(TInflator* s);
int    jdb_inflator_take_leftover(TInflator* s, uint8* dst);

#endif
