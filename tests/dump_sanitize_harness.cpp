// ASan/UBSan harness for the -dump formatter (test infrastructure)
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <vector>
#include <string>
#include "../include/merfin_b200.h"
void mfn_set_error(const char *fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); }
int main() {
  const uint64_t n = 700000;
  std::vector<uint32_t> vr(n), va(n); std::vector<uint8_t> ok(n);
  uint64_t s = 12345;
  const uint32_t R[] = {0, 1, 3, 14, 15, 16, 29, 2047, 2048, 5000, 4000000000u}, A[] = {0, 1, 2, 3, 31, 32, 70000};
  for (uint64_t i = 0; i < n; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; vr[i] = R[(s >> 33) % 11]; va[i] = A[(s >> 40) % 7]; ok[i] = ((s >> 50) % 50) != 0; }
  const char *names[] = {"s", "contig_with_a_name_longer_than_thirty_two_characters_0001", "xxxxxxxxxxxxxxxxxxxxxxxxxxxxxxxx"};
  double peaks[] = {15.0, 1e-7, 1e-300};
  int fd = open("/dev/null", O_WRONLY);
  for (int m = 0; m < 3; m++) for (int t : {1, 3, 8}) for (const char *nm : names) {
    mfn_kmodel km = {peaks[m], 0, nullptr, nullptr};
    mfn_seq_counts c;
    // exact-size heap copies so that any over-read of the inputs is caught
    std::vector<uint32_t> a(vr.begin(), vr.end()), b(va.begin(), va.end()); std::vector<uint8_t> o(ok.begin(), ok.end());
    std::string ident(nm);
    char *idc = (char *)malloc(ident.size() + 1); memcpy(idc, ident.c_str(), ident.size() + 1);
    int st = mfn_dump_write(fd, idc, n, a.data(), b.data(), o.data(), &km, t, &c);
    free(idc);
    if (st) { printf("status %d\n", st); return 1; }
  }
  printf("ok\n");
  return 0;
}
