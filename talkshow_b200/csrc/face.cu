// talkshow_b200 — face regressor (s2g_face.Generator): placeholder translation unit until the
// wav2vec2 path lands; ts_load_face / ts_face_forward report TS_ERR_UNSUPPORTED.
#include "pixelcnn.h"
namespace ts {
void face_destroy(ts_engine*) {}
}
extern "C" int ts_load_face(ts_engine* e, const ts_tensor*, int) {
  if (!e) return TS_ERR_INVALID;
  e->err = "face path not built yet";
  return TS_ERR_UNSUPPORTED;
}
extern "C" int ts_face_forward(ts_engine* e, const float*, const float*, float*, int, int, int, void*) {
  if (!e) return TS_ERR_INVALID;
  e->err = "face path not built yet";
  return TS_ERR_UNSUPPORTED;
}
