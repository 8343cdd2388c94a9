/* oracle/ref_getiq_glue.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Harness around the reference's own 18-bit codec, assembled from getiq64.s (see nasm2gas.py):
 *   getiq_check expand   N  < packed bytes (9 per 4 words)   > N*4 words' bytes (expand_rawdat, getiq64.s:158-220)
 *   getiq_check compress N  < N words (int32)                > 9*N/4 bytes      (compress_rawdat_disk, getiq64.s:98-156)
 * The routines work on Linrad's globals (rx_read_bytes bytes of timf1_char at timf1p_pa / timf1p_pc_disk and the
 * rawsave_tmp buffers); expand_rawdat reads two bytes before and one byte behind the packed block, as it does in
 * Linrad's larger buffers, hence the padding. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

char *rawsave_tmp, *rawsave_tmp_net, *rawsave_tmp_disk, *fft1_char, *timf1_char;
int rx_read_bytes, timf1p_pa, timf1p_pc_disk, timf1p_pc_net;
void expand_rawdat(void);
void compress_rawdat_disk(void);
void compress_rawdat_net(void);

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  const long nwords = atol(argv[2]);
  if (nwords <= 0 || nwords % 4) return 2;
  const size_t wbytes = (size_t)nwords * 4, pbytes = (size_t)nwords / 4 * 9;
  char *words = calloc(wbytes + 64, 1), *packed = calloc(pbytes + 64, 1);
  if (!words || !packed) return 3;
  timf1_char = words + 32;
  rawsave_tmp = rawsave_tmp_disk = rawsave_tmp_net = packed + 32;
  timf1p_pa = timf1p_pc_disk = timf1p_pc_net = 0;
  rx_read_bytes = (int)wbytes;
  if (!strcmp(argv[1], "expand")) {
    if (fread(packed + 32, 1, pbytes, stdin) != pbytes) return 4;
    expand_rawdat();
    fwrite(words + 32, 1, wbytes, stdout);
  } else if (!strcmp(argv[1], "compress") || !strcmp(argv[1], "compress_net")) {
    if (fread(words + 32, 1, wbytes, stdin) != wbytes) return 4;
    if (argv[1][8]) compress_rawdat_net(); else compress_rawdat_disk();
    fwrite(packed + 32, 1, pbytes, stdout);
  } else return 2;
  return 0;
}
