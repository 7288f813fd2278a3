/* mrg_restart.h -- the particle records of the reference's restart file, written straight from HBM.
 *
 * restrt(iresrt=2) (F:9522-9856) gathers the strided ownership of all ranks with 12 mpi_allreduce of np0 doubles into 12
 * scratch arrays (F:9622-9668) and lets rank 0 write unit 12 with `form='unformatted'` (F:9696-9728).  Its last four
 * records are the particles:
 *     write(12) qmulti,wmulti,qmulte,wmulte           F:9722
 *     write(12) npr                                    F:9723
 *     write(12) xxi,yyi,zzi,vvxi,vvyi,vvzi             F:9724   six arrays of np0 doubles, one record
 *     write(12) xxe,yye,zze,vvxe,vvye,vvze             F:9725
 * With the particles resident on the GPU these four records are produced here in the same byte format (Fortran
 * unformatted sequential access as gfortran writes it: every record framed by two 4-byte length markers, records longer
 * than 2^31-9 bytes split into subrecords whose markers carry the continuation in their sign), in the original l order,
 * streaming through one pinned chunk instead of 12 host arrays.  The host's restrt keeps writing records 1-8 (scalars,
 * fields, moments, plot averages: not on this path) and appends these.  Single-rank contexts write the whole file
 * section; a rank of a multi-rank job writes its owned subset into its own file (first / stride are stored in a trailing
 * record of this library, not of the reference) -- merging them is host work.
 *
 * PARITY: pinned to the reference's own restrt.  The image has no Fortran compiler, so the reference's restrt runs through
 * the translated code (oracle/f03c.py turns its write(12)/read(12) statements into record calls of oracle/ref_runtime.c,
 * an independent implementation of the same framing).  tests/test_restart_records.py: the file that restrt(iresrt=2)
 * writes after a step of the reference's time cycle holds 12 records; the last four are byte-identical to what this
 * library writes for the same particles -- from host arrays and from particles resident (and sorted) in HBM -- and the
 * reference's restrt(iresrt=1) reads a file whose particle records came from here.  What stays a convention is the
 * marker format itself (gfortran's documented one); both implementations of it agree byte for byte, sub-records included.
 */
#ifndef MRG_RESTART_H
#define MRG_RESTART_H
#include <stdint.h>
#include "../../include/mrg_fulmov.h"
#ifdef __cplusplus
extern "C" {
#endif
/* One Fortran unformatted sequential record made of `nparts` pieces (a Fortran I/O list), appended to / read from an
 * open stdio stream.  Returns 0 on success.                                                                         */
/* gfortran's -fmax-subrecord-length (default 2^31-9 bytes; 0 restores it): tests use a small value to exercise the split. */
void mrg_f77_set_max_subrecord(uint64_t bytes);
int mrg_f77_write_record(void* file, const void* const* parts, const uint64_t* bytes, int32_t nparts);
int mrg_f77_read_record(void* file, void* const* parts, const uint64_t* bytes, int32_t nparts);
/* Records F:9722-9725 for species 1 (ions) and 2 (electrons) of `ctx`, appended to `path`; np0 = declared array length
 * (param_080A.h: entries beyond npr are written as zeros, as the reference's zeroed scratch arrays are).            */
int mrg_restart_append_particles(mrg_ctx* ctx, const char* path, double qmulti, double wmulti, double qmulte, double wmulte,
                                 int64_t npr, int64_t np0, int64_t first, int64_t stride);
/* The inverse: skips `skip_records` records of `path` (the host's records 1-8), reads the four particle records and
 * uploads the owned subset (first, stride) of both species into `ctx`.  Outputs the scalars.                        */
int mrg_restart_read_particles(mrg_ctx* ctx, const char* path, int32_t skip_records, double qw[4], int64_t* npr,
                               int64_t np0, int64_t first, int64_t stride);
#ifdef __cplusplus
}
#endif
#endif
