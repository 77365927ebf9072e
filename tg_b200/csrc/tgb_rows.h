/*
 * tgb_rows.h -- which rows of the frame a rank shades, and where a row lives in the frame buffers.
 *
 * GI rays are split across GPUs by screen tile. A tile of CONTIGUOUS rows gives one rank the sky and
 * another all the hit pixels (measured at N = 8: the top tile's shading took 0.06 ms, the frame waited 0.45 ms for the
 * others), so the frame is cut into bands of 16 rows -- the height of a k_shade CTA -- and band b belongs to rank b mod N.
 * The device buffers (visibility, material words, radiance, presented frame) store rows in VIRTUAL order, rank-major:
 *
 *     virtual row v = rank * tile_rows + (b / N) * 16 + (p mod 16)        for physical row p, band b = p / 16, rank = b mod N
 *
 * so that a rank's rows are one contiguous range [rank * tile_rows, (rank + 1) * tile_rows) and every exchange (reduce-scatter,
 * peer-memory merge, all-gather, the frame sink's band copies) stays a contiguous block. tile_rows = 16 * ceil(ceil(H / 16) / N);
 * virtual rows whose physical row is >= H are padding. With one rank the mapping is the identity. Only three places see
 * physical rows: K1's resolve (pixel -> buffer index), k_shade (buffer row -> pixel, for the ray and the RNG seed) and the host
 * read-back / write paths.
 */
#ifndef TGB_ROWS_H
#define TGB_ROWS_H

#include "tgb_math.h"

#define TGB_BAND_ROWS 16u

TGB_HD u32 tgb_tile_rows_for(u32 height, u32 n_ranks)
{
    const u32 n_bands = (height + TGB_BAND_ROWS - 1u) / TGB_BAND_ROWS;
    return ((n_bands + n_ranks - 1u) / n_ranks) * TGB_BAND_ROWS;
}

TGB_HD u32 tgb_row_to_virtual(u32 physical_row, u32 n_ranks, u32 tile_rows)
{
    const u32 band = physical_row / TGB_BAND_ROWS;
    return (band % n_ranks) * tile_rows + (band / n_ranks) * TGB_BAND_ROWS + (physical_row % TGB_BAND_ROWS);
}

/* may be >= height: a padding row */
TGB_HD u32 tgb_row_to_physical(u32 virtual_row, u32 n_ranks, u32 tile_rows)
{
    const u32 rank = virtual_row / tile_rows, band_in_tile = (virtual_row % tile_rows) / TGB_BAND_ROWS;
    return (band_in_tile * n_ranks + rank) * TGB_BAND_ROWS + (virtual_row % TGB_BAND_ROWS);
}

#endif
