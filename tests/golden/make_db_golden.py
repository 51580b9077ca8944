"""Generate tests/golden/db_postcard.json: postcard byte vectors for the three metadata files of an
elodin-db directory (schema / metadata / db_state), produced by the REFERENCE's own C implementation of
the wire format (/root/reference/libs/postcard-c/postcard.h, the header `elodin-db gen-cpp` ships to C++
clients).  The driver below is ours; the header is compiled where it lies and never copied.  Run in the
build container only (the GPU box has no /root/reference); the JSON is what travels.

    python tests/golden/make_db_golden.py
"""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/libs/postcard-c"

DRIVER = r'''
#include "postcard.h"
#include <stdio.h>
#include <string.h>
static uint8_t buf[4096];
static postcard_slice_t s;
static void begin(void) { postcard_init_slice(&s, buf, sizeof buf); }
static void end(const char *name, int last) {
    printf("  \"%s\": \"", name);
    for (size_t i = 0; i < s.len; ++i) printf("%02x", s.data[i]);
    printf("\"%s\n", last ? "" : ",");
}
static void str(const char *v) { postcard_encode_string(&s, v, strlen(v)); }
int main(void) {
    printf("{\n");
    /* Schema{prim_type: PrimType, shape: Vec<u64>} */
    begin(); postcard_encode_variant(&s, 10); postcard_start_seq(&s, 1); postcard_encode_u64(&s, 7); end("schema_f64_7", 0);
    begin(); postcard_encode_variant(&s, 10); postcard_start_seq(&s, 0); end("schema_f64_scalar", 0);
    begin(); postcard_encode_variant(&s, 3); postcard_start_seq(&s, 2); postcard_encode_u64(&s, 2); postcard_encode_u64(&s, 300); end("schema_u64_2x300", 0);
    /* ComponentMetadata{component_id: ComponentId(u64), name: String, metadata: HashMap<String,String>} */
    begin(); postcard_encode_u64(&s, 6412982479418929775ULL); str("a.world_pos"); postcard_start_map(&s, 1); str("priority"); str("5"); end("metadata_a_world_pos", 0);
    begin(); postcard_encode_u64(&s, 0x7fffffffffffffffULL); str("\xcf\x89x"); postcard_start_map(&s, 0); end("metadata_max_id_utf8", 0);
    /* DbConfig{recording: bool, default_stream_time_step: Duration{secs: u64, nanos: u32}, metadata} */
    begin(); postcard_encode_bool(&s, false); postcard_encode_u64(&s, 0); postcard_encode_u32(&s, 16666667); postcard_start_map(&s, 1); str("time.start_timestamp"); str("1767225600000000"); end("db_state_60hz", 0);
    begin(); postcard_encode_bool(&s, true); postcard_encode_u64(&s, 2); postcard_encode_u32(&s, 500000000); postcard_start_map(&s, 0); end("db_state_recording_2500ms", 0);
    /* scalars */
    begin(); postcard_encode_i64(&s, -1); end("i64_minus_1", 0);
    begin(); postcard_encode_i64(&s, 1767225600000000LL); end("i64_timestamp", 0);
    begin(); postcard_encode_u64(&s, 18446744073709551615ULL); end("u64_max", 1);
    printf("}\n");
    return 0;
}
'''

with tempfile.TemporaryDirectory() as d:
    src = os.path.join(d, "driver.cpp")
    open(src, "w").write(DRIVER)
    exe = os.path.join(d, "driver")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", REF, src, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
vectors = json.loads(out)
json.dump({"_source": "reference libs/postcard-c/postcard.h compiled by tests/golden/make_db_golden.py", **vectors},
          open(os.path.join(HERE, "db_postcard.json"), "w"), indent=1)
print(json.dumps(vectors, indent=1))
