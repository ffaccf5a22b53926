/* stand-in for VTK's vtkFiltersCoreModule.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
