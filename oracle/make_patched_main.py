"""Reference-side proof of the drop-in boundary (test infrastructure; build output only, never product code).

    python oracle/make_patched_main.py [/root/reference] [oracle/_ref]

Writes <out>/mc_main_gpu.cc: a build-time COPY of the reference's own mc_main.cc with the patch of INTEGRATION.md section B
applied mechanically -- its input parsing, set-up, block loop, Save* writers and checkpoint code stay as they are; the
body of the `time` loop (PIMCPass) and MCGetAverage are routed through the C ABI of libpimcgpu.so.  oracle/Makefile
(`make ref_gpu`) compiles it against the reference's unmodified translation units and links libpimcgpu.so into
<out>/pimc_ref_gpu.  tests/test_gpu_boundary.py runs that binary next to pimc_b200 on the reference's CO2 deck.
Nothing of the reference is committed: the copy lives under oracle/_ref (git-ignored)."""
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
src = open(os.path.join(ref, "mc_main.cc")).read()

GLUE = r'''
// ---- glue inserted by oracle/make_patched_main.py (INTEGRATION.md section B) -------------------------------------------
#include <cstring>
#include "pimcgpu.h"
extern double *_pgrid1D[], *_poten1D[]; extern int _psize1D[];                                  // mc_poten.cc:17-21
extern double *_rgrid2D[], *_cgrid2D[], **_poten2D[]; extern int _rsize2D[], _csize2D[]; extern double _delta_r[], _delta_c[];
extern double *_rotgrid[], *_rotdens[], *_rotderv[], *_rotesqr[]; extern int _rotsize[];
extern double **_rcf;
static void gpu_die(const char *who) { nrerror(who, pimcgpu_last_error()); }
static void gpu_attach(void)
{
   if (WORM) nrerror("gpu_attach", "this proof-of-boundary build routes the non-worm block loop only");
   pimcgpu_system s; pimcgpu_tables t;
   memset(&s, 0, sizeof s); memset(&t, 0, sizeof t);
   s.ntypes = NumbTypes;  s.P = NumbTimes;  s.Q = ROTATION ? NumbRotTimes : 0;
   s.temperature = Temperature;  s.ispher = ISPHER;  s.minimage = MINIMAGE;
   s.rotden_type = RotDenType; s.rot_odevn = RotOdEvn; s.rot_eoff = RotEoff; s.x_rot = X_Rot; s.y_rot = Y_Rot; s.z_rot = Z_Rot; s.rnratio = RNratio;
   s.reflect[0] = IREFLX; s.reflect[1] = IREFLY; s.reflect[2] = IREFLZ; s.rotsym = IROTSYM; s.nfold_rot = NFOLD_ROT;
   for (int d = 0; d < 3; d++) s.box[d] = BoxSize[d];
   for (int k = 0; k < NumbTypes; k++) {
      s.type[k].numb = MCAtom[k].numb;   s.type[k].molecule = MCAtom[k].molecule;  s.type[k].stat = MCAtom[k].stat;
      s.type[k].levels = MCAtom[k].levels; s.type[k].mass = MCAtom[k].mass;
      s.type[k].mcstep = MCAtom[k].mcstep; s.type[k].rtstep = MCAtom[k].rtstep;
      if (MCAtom[k].molecule == 0) { t.n1d = _psize1D[k]; t.grid1d = _pgrid1D[k]; t.pot1d = _poten1D[k]; }
      if (MCAtom[k].molecule == 1) {
         if (_rsize2D[k] > 1) {
            t.rsize2d = _rsize2D[k]; t.csize2d = _csize2D[k]; t.dr2d = _delta_r[k]; t.dc2d = _delta_c[k];
            t.rgrid2d = _rgrid2D[k]; t.cgrid2d = _cgrid2D[k]; t.pot2d = _poten2D[k][0];
         }
         if (ROTATION && RotDenType == 0) { t.nrot = _rotsize[k]; t.rotgrid = _rotgrid[k]; t.rotdens = _rotdens[k]; t.rotderv = _rotderv[k]; t.rotesqr = _rotesqr[k]; }
      }
   }
   if (IMTYPE >= 0 && MCAtom[IMTYPE].molecule == 2) {
      if (NumbTypes > 1) { t.rgrd = Rgrd; t.thgrd = THgrd; t.chgrd = CHgrd; t.rvmin = Rvmin; t.rvmax = Rvmax; t.vtable = vtable; }
      if (ROTATION && RotDenType == 0) { t.rho3d = rhoprp; t.erot3d = erotpr; t.esq3d = erotsq; }
   }
   s.nchains = 1;
   if (pimcgpu_init(&s, &t)) gpu_die("pimcgpu_init");
   // MCCoords / MCAngles are doubleMatrix blocks: row 0 points at [3][N*P] contiguous doubles (mc_utils.cc:10-30)
   if (pimcgpu_upload_state(-1, MCCoords[0], MCAngles[0], BOSONS ? PIndex : NULL)) gpu_die("pimcgpu_upload_state");
   unsigned long seed[6] = {12345, 12345, 12345, 12345, 12345, 12345};      // fixedseed(), omprng.cc:14-18
   if (pimcgpu_seed(seed)) gpu_die("pimcgpu_seed");
}
static void gpu_pass(int type) { if (type == 0 && pimcgpu_steps(1)) gpu_die("pimcgpu_steps"); }     // PIMCPass(type, time) of every type
static void gpu_download(void)
{
   if (pimcgpu_download_state(0, MCCoords[0], MCAngles[0], MCCosine[0], BOSONS ? PIndex : NULL)) gpu_die("pimcgpu_download_state");
}
static void gpu_block_begin(void) { if (pimcgpu_accum_reset()) gpu_die("pimcgpu_accum_reset"); }
static void gpu_get_average(void)                                                                    // the whole MCGetAverage on the device
{
   avergCount += 1.0;  totalCount += 1.0;
   if (pimcgpu_measure()) gpu_die("pimcgpu_measure");
   if (PrintXYZprl) gpu_download();
}
static void gpu_counters(void)
{
   pimcgpu_scalars b;
   pimcgpu_accum_device_ptr();
   if (pimcgpu_sync() || pimcgpu_block_scalars(&b)) gpu_die("pimcgpu_block_scalars");
   for (int k = 0; k < NumbTypes; k++) for (int m = 0; m < 3; m++) { MCTotal[k][m] = b.mctotal[k][m]; MCAccep[k][m] = b.mcaccep[k][m]; }
}
static void gpu_block_end(void)
{
   pimcgpu_scalars b;
   pimcgpu_accum_device_ptr();
   if (pimcgpu_sync() || pimcgpu_block_scalars(&b)) gpu_die("pimcgpu_block_scalars");
   _bkin = b.kin; _bpot = b.pot; _brot = b.rot; _brotsq = b.rotsq; _bCv = b.cv; _bCv_trans = b.cv_trans; _bCv_rot = b.cv_rot;
   _kin_total += b.kin; _pot_total += b.pot; _rot_total += b.rot; _rotsq_total += b.rotsq; _Cv_total += b.cv; _Cv_trans_total += b.cv_trans; _Cv_rot_total += b.cv_rot;
   for (int k = 0; k < NumbTypes; k++) for (int m = 0; m < 3; m++) { MCTotal[k][m] = b.mctotal[k][m]; MCAccep[k][m] = b.mcaccep[k][m]; }
   if (ROTATION) {
      long n = 0, off_rcf = 0;
      pimcgpu_accum_layout(&n, NULL, NULL, NULL, NULL, &off_rcf, NULL);
      double *acc = new double[n];
      if (pimcgpu_accum_download(acc, n)) gpu_die("pimcgpu_accum_download");
      for (int it = 0; it < NumbRotTimes; it++) _rcf[0][it] = acc[off_rcf + it];
      delete[] acc;
   }
}
// ---- end of glue ----------------------------------------------------------------------------------------------------
'''


def sub(pattern, repl, text, count=1, flags=re.M):
    new, n = re.subn(pattern, repl, text, count=count, flags=flags)
    assert n >= 1, f"patch anchor not found: {pattern}"
    return new


i = src.index("int main(")
patched = src[:i] + GLUE + src[i:]
patched = sub(r"^(\s*)randomseed\(\);", r"\1randomseed(); gpu_attach();", patched)
patched = sub(r"^(\s*)else\s*\n\s*PIMCPass\(type,time\);", r"\1else\n\1gpu_pass(type);", patched)
patched = sub(r"^(\s*)MCResetBlockAverage\(\);\s*\n(\s*)\n(\s*)long int passCount = 0;", r"\1MCResetBlockAverage(); gpu_block_begin();\n\3long int passCount = 0;", patched)
head, body = patched[:patched.index("int main(")], patched[patched.index("int main("):]
end_main = body.index("\nvoid PIMCPass")
main_txt, rest = body[:end_main], body[end_main:]
main_txt, n = re.subn(r"^(\s+)MCGetAverage\(\);", r"\1gpu_get_average();", main_txt, flags=re.M)
assert n == 2, n
main_txt = sub(r"^(\s*)if \(blockCount>NumberOfEQBlocks && avergCount\)   // skip equilibration steps", r"\1gpu_block_end();\n\1if (blockCount>NumberOfEQBlocks && avergCount)", main_txt)
main_txt = sub(r"^(\s*)IOFileBackUp\(FSTATUS\); StatusIO\(IOWrite,FSTATUS\);", r"\1gpu_download();\n\1IOFileBackUp(FSTATUS); StatusIO(IOWrite,FSTATUS);", main_txt)
main_txt = sub(r"^(\s*)MCSaveAcceptRatio\(passTotal,passCount,blockCount\);", r"\1{ gpu_counters(); MCSaveAcceptRatio(passTotal,passCount,blockCount); }", main_txt)
patched = head + main_txt + rest
os.makedirs(out, exist_ok=True)
open(os.path.join(out, "mc_main_gpu.cc"), "w").write(patched)
print("wrote", os.path.join(out, "mc_main_gpu.cc"))
