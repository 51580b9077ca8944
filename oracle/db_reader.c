/* TEST INFRASTRUCTURE — not part of the product, never linked into libb200_sixdof.so.
 *
 * A plain-C restatement of the READER side of the reference's elodin-db time series, so that the directories
 * elodin_b200/db_sink.py writes are opened by code that follows the reference's own read path instead of by the
 * writer's author:
 *
 *   libs/db/src/append_log.rs:46-52    struct Header { committed_len: u64, head_len: u64, extra: E }   (repr(C), 24 bytes)
 *   libs/db/src/append_log.rs:84-96    AppendLog::open  — map the file, nothing else
 *   libs/db/src/append_log.rs:131-143  len() = committed_len - size_of(Header); data() = bytes [24, committed_len)
 *   libs/db/src/time_series.rs:40-54   TimeSeries::open — <dir>/index (extra = start Timestamp, i64 us) and <dir>/data
 *                                      (extra = element size in bytes, u64)
 *   libs/db/src/time_series.rs:56-62   start_timestamp() = min(index extra, first timestamp)
 *   libs/db/src/time_series.rs:75-82   timestamps() = the index log's data as [i64]; element_size() = data extra
 *   libs/db/src/time_series.rs:84-90   get(ts): exact binary search, row = data[i * es, (i + 1) * es)
 *   libs/db/src/time_series.rs:98-110  get_nearest(ts): Ok(i) -> i, Err(i) -> i - 1 (saturating); None when out of range
 *   libs/db/src/time_series.rs:131-139 range_indices(start..end): [partition_point(t < start), partition_point(t <= end));
 *                                      None when empty — the END of the range is inclusive
 *
 * The reference maps the whole sparse 8 GiB file; this reader reads the header and the committed bytes only (the
 * sparse tail is zeros by construction and never addressed by the functions above).  A row count that does not match
 * between the two logs is reported, not repaired: rows = min(index rows, data rows), as every accessor above would
 * effectively see it (an index entry whose row is missing makes `data.get(range)` return None).
 *
 * Parity status: restated from source, exercised against directories written by db_sink.py (tests/test_db_sink.py);
 * the reference's binary cannot be built here (Rust), so this is "reader restated", not "reader run". */
#define _FILE_OFFSET_BITS 64
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LOG_HEADER 24u

typedef struct {
    uint64_t committed_len, head_len;
    uint8_t extra[8];
    uint8_t *data;     /* committed bytes behind the header */
    uint64_t len;      /* committed_len - 24 */
} orc_append_log;

typedef struct {
    orc_append_log index, data;
    uint64_t rows;          /* min(index rows, data rows) */
    uint64_t element_size;
    int64_t index_extra;    /* start timestamp written at creation */
} orc_series;

static int log_open(const char *path, orc_append_log *out)
{
    memset(out, 0, sizeof *out);
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    uint8_t h[ORC_LOG_HEADER];
    if (fread(h, 1, sizeof h, f) != sizeof h) { fclose(f); return -2; }
    memcpy(&out->committed_len, h, 8);      /* little-endian host, like the reference's targets */
    memcpy(&out->head_len, h + 8, 8);
    memcpy(out->extra, h + 16, 8);
    if (out->committed_len < ORC_LOG_HEADER) { fclose(f); return -3; }
    out->len = out->committed_len - ORC_LOG_HEADER;
    out->data = (uint8_t *)malloc(out->len ? out->len : 1);
    if (!out->data) { fclose(f); return -4; }
    if (out->len && fread(out->data, 1, out->len, f) != out->len) { free(out->data); out->data = 0; fclose(f); return -5; } /* file shorter than it claims */
    fclose(f);
    return 0;
}

int orc_series_open(const char *dir, orc_series *s)
{
    char p[4096];
    memset(s, 0, sizeof *s);
    if (snprintf(p, sizeof p, "%s/index", dir) >= (int)sizeof p) return -10;
    int rc = log_open(p, &s->index);
    if (rc) return rc;
    snprintf(p, sizeof p, "%s/data", dir);
    rc = log_open(p, &s->data);
    if (rc) { free(s->index.data); s->index.data = 0; return rc - 100; }
    memcpy(&s->index_extra, s->index.extra, 8);
    memcpy(&s->element_size, s->data.extra, 8);
    const uint64_t irows = s->index.len / 8;
    const uint64_t drows = s->element_size ? s->data.len / s->element_size : 0;
    s->rows = irows < drows ? irows : drows;
    return 0;
}

void orc_series_close(orc_series *s)
{
    free(s->index.data);
    free(s->data.data);
    memset(s, 0, sizeof *s);
}

uint64_t orc_series_rows(const orc_series *s) { return s->rows; }
uint64_t orc_series_index_rows(const orc_series *s) { return s->index.len / 8; }
uint64_t orc_series_element_size(const orc_series *s) { return s->element_size; }
int64_t orc_series_index_extra(const orc_series *s) { return s->index_extra; }
const int64_t *orc_series_timestamps(const orc_series *s) { return (const int64_t *)s->index.data; }
const uint8_t *orc_series_data(const orc_series *s) { return s->data.data; }

int64_t orc_series_start_timestamp(const orc_series *s)
{
    const int64_t *ts = (const int64_t *)s->index.data;
    if (s->index.len >= 8 && ts[0] < s->index_extra) return ts[0];
    return s->index_extra;
}

/* slice::binary_search: 1 and *pos = index when found, 0 and *pos = insertion point otherwise */
static int bsearch_ts(const int64_t *ts, uint64_t n, int64_t key, uint64_t *pos)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (ts[mid] < key) lo = mid + 1;
        else if (ts[mid] > key) hi = mid;
        else { *pos = mid; return 1; }
    }
    *pos = lo;
    return 0;
}

/* TimeSeries::get: the row stamped exactly `timestamp`, or -1 */
int64_t orc_series_get(const orc_series *s, int64_t timestamp)
{
    uint64_t i;
    if (!bsearch_ts((const int64_t *)s->index.data, s->index.len / 8, timestamp, &i)) return -1;
    if ((i + 1) * s->element_size > s->data.len) return -1; /* data.get(range) == None */
    return (int64_t)i;
}

/* TimeSeries::get_nearest: the row at or before `timestamp` (the first row when it precedes them all), or -1 */
int64_t orc_series_get_nearest(const orc_series *s, int64_t timestamp)
{
    const uint64_t n = s->index.len / 8;
    uint64_t i;
    if (!bsearch_ts((const int64_t *)s->index.data, n, timestamp, &i)) i = i ? i - 1 : 0;
    if (i >= n) return -1;
    if ((i + 1) * s->element_size > s->data.len) return -1;
    return (int64_t)i;
}

/* TimeSeries::range_indices: rows with start <= t <= end as [*i0, *i1); 0 when there are none */
int orc_series_range(const orc_series *s, int64_t start, int64_t end, uint64_t *i0, uint64_t *i1)
{
    const int64_t *ts = (const int64_t *)s->index.data;
    const uint64_t n = s->index.len / 8;
    uint64_t a = 0, b = 0;
    while (a < n && ts[a] < start) ++a;   /* partition_point(|t| t < start); the logs are sorted (push rejects time travel) */
    b = a;
    while (b < n && ts[b] <= end) ++b;    /* partition_point(|t| t <= end) */
    if (a >= b) return 0;
    *i0 = a; *i1 = b;
    return 1;
}
