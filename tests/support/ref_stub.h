// Minimal stand-ins for the reference types the INTEGRATION.md adapter touches (names and members as in data_struct.h:788-886,
// 1255-1346); only what the snippet reads.  Used by tests/test_abi.py to compile the snippet against include/graspa_b200.h.

#include <vector>
#include <stdexcept>
#include <algorithm>
#include <cstdint>
#include <cstddef>
struct double3 { double x, y, z; }; struct int3 { int x, y, z; };
struct Atoms { double3* pos; double* scale; double* charge; double* scaleCoul; size_t* Type; size_t* MolID; size_t size, Molsize, Allocate_size; };
struct ForceField { double* epsilon; double* sigma; double* z; double* shift; double* C10; double CutOffVDW, CutOffCoul, OverlapCriteria; size_t size; bool noCharges, VDWRealBias, Use1264; };
struct Tail { bool UseTail; double Energy; };
struct Boxsize { double* Cell; double* InverseCell; bool Cubic; double Volume, Alpha, Prefactor, ReciprocalCutOff; int3 kmax; };
struct WidomStruct { size_t NumberWidomTrials, NumberWidomTrialsOrientations; };
struct RandomNumber { double3* host_random; size_t randomsize; };
struct Components { int3 NComponents; std::vector<Atoms> HostSystem; std::vector<double> ExclusionIntra, ExclusionAtom; std::vector<bool> rigid, hasPartialCharge; double Beta; std::vector<Tail> TailCorrection; };
struct Variables { std::vector<Components> SystemComponents; std::vector<Boxsize> Box; ForceField FF; std::vector<WidomStruct> Widom; RandomNumber Random; };
