/* stand-in for VTK's vtkObjectFactory.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
