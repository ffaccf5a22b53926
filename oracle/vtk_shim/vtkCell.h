/* stand-in for VTK's vtkCell.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
